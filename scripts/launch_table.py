"""Per-kernel totals of an ncu --metrics gpu__time_duration.sum --csv launch list: launch_table.py launches.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")) if len(r) > 5]
hdr = rows[0]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
iu = hdr.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
    name = r[ik].split("(")[0][:70]
    tot[name][0] += 1
    tot[name][1] += v
all_us = sum(v[1] for v in tot.values())
print(f"{'kernel':70s} {'launches':>9s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:70s} {n:9d} {us:12.1f} {us / n:10.2f} {100 * us / all_us:6.1f}%")
