#!/usr/bin/env python
"""Static facts about every kernel in libbayesssm_b200.so, read off the cubin here (no GPU): registers, static shared
memory, local (spill) bytes from `cuobjdump -res-usage`, and counts of the SASS mnemonics that matter for this
engine from `cuobjdump -sass` (LDGSTS = cp.async staging, MUFU = SFU ops of Box-Muller / sin / exp, DFMA/DADD/DMUL = the
fp64 accumulation, ATOM/RED = atomics, BAR = block barriers, SHFL = warp shuffles, STL/LDL = spills).
usage: python scripts/sass_resources.py > profiles/r1_sass_resources.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bayesssm_b200", "libbayesssm_b200.so")
MNEMONICS = ["LDGSTS", "LDG", "STG", "LDS", "STS", "MUFU", "FFMA", "DFMA", "DADD", "DMUL", "IMAD", "ATOM", "RED", "BAR", "SHFL", "STL", "LDL"]


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.strip().splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"\(.*\)$", "", name)
    name = name.replace("bssm::", "").replace("void ", "")
    return name


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
        usage[m.group(1)] = tuple(int(x) for x in m.groups()[1:])
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = dict.fromkeys(MNEMONICS, 0)
            counts[cur]["total"] = 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["total"] += 1
            for k in MNEMONICS:
                if op == k or (k in ("ATOM", "RED", "BAR", "SHFL") and op.startswith(k)) or (k == "ATOM" and op.startswith("ATOMG")):
                    counts[cur][k] += 1
                    break
    names = demangle(sorted(usage))
    print("arch: sm_100a; library:", os.path.relpath(LIB, ROOT))
    print(f"{'kernel':78s} {'REG':>4s} {'SMEM':>6s} {'STACK':>5s} {'LOCAL':>5s} {'SASS':>6s} " + " ".join(f"{k:>6s}" for k in MNEMONICS))
    for mangled in sorted(usage, key=lambda k: short(names[k])):
        reg, stack, shared, local = usage[mangled]
        c = counts.get(mangled, {})
        print(f"{short(names[mangled])[:78]:78s} {reg:4d} {shared:6d} {stack:5d} {local:5d} {c.get('total', 0):6d} "
              + " ".join(f"{c.get(k, 0):6d}" for k in MNEMONICS))
    spills = [short(names[k]) for k, v in usage.items() if v[3] > 0 or counts.get(k, {}).get("STL", 0) > 0]
    print("\nkernels with local memory (spills or local arrays):", ", ".join(spills) if spills else "none")


if __name__ == "__main__":
    sys.exit(main())
