#!/bin/bash
# A/B: chain-persistent kernel on / off (in-tree library), PMMH 1024 / 256 / 128 chains
cd /root/repo
for ch in 0 1; do
  export BSSM_ST_CHAIN=$ch
  for C in 1024 256 128; do
    r=$(python bench.py --workload pmmh --chains $C --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'])")
    echo "chain=$ch chains=$C ms/iter it/s frac: $r" | tee -a gpurun_out/ab_chain.txt
  done
done
