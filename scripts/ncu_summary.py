"""Summarise an ncu report (--set full, --import-source on) of the persistent kernel into text:
key metrics, instruction mix by opcode, hottest source lines.  Usage: ncu_summary.py report.ncu-rep"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
print("== key metrics (one launch) ==")
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for k in KEYS:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k:70s} {vals[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]
ia, isrc = h.index("Instructions Executed"), h.index("Source")
ops, tot = collections.Counter(), 0
for r in rows[2:]:
    try:
        n = int(r[ia])
    except Exception:
        continue
    s = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip())
    ops[(s.split()[0] if s else "?").split(".")[0]] += n
    tot += n
print(f"\n== warp instructions executed: {tot} ==")
for op, n in ops.most_common(24):
    print(f"{op:10s} {100 * n / tot:5.1f}%")
cs = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(cs.splitlines()))
cur, agg, ti, ts = None, [], 0, 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name", ""):
        continue
    try:
        line, samp, inst = int(r[0]), int(r[4]), int(r[7])
    except Exception:
        continue
    agg.append((cur, line, r[1].strip()[:96], samp, inst))
    ti += inst
    ts += samp
print("\n== hottest source lines by stall samples (share of samples | share of instructions) ==")
for a in sorted(agg, key=lambda a: -a[3])[:16]:
    print(f"{a[0]}:{a[1]:<4d} {100 * a[3] / ts:5.1f}% | {100 * a[4] / ti:5.1f}%  {a[2]}")
print("\n== hottest source lines by instructions ==")
for a in sorted(agg, key=lambda a: -a[4])[:16]:
    print(f"{a[0]}:{a[1]:<4d} {100 * a[4] / ti:5.1f}% | {100 * a[3] / ts:5.1f}%  {a[2]}")
