#!/bin/bash
# A/B over variant libraries: pmmh 1024 / 128 chains, single filter N=2^24
cd /root/repo
for v in "$@"; do
  export BSSM_LIB_PATH=/root/repo/variants/lib_$v.so
  a=$(python bench.py --workload pmmh --chains 1024 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])")
  b=$(python bench.py --workload pmmh --chains 128 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])")
  c=$(python bench.py --engine stream --N 16777216 --T 200 --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value']/1e9, d['roofline']['frac'])")
  d=$(python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value']/1e9)")
  echo "$v | pmmh1024 $a | pmmh128 $b | N2p24 $c | C2 $d" | tee -a gpurun_out/ab_stream.txt
done
