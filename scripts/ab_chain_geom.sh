#!/bin/bash
# chain-persistent kernel: blocks per filter / block size sweep at 128 and 256 chains
cd /root/repo
export BSSM_ST_CHAIN=1
for C in 128 256; do
for th in 128 256; do
  for b in 2 4 6 9 16; do
    export BSSM_ST_THREADS=$th BSSM_ST_BPC=$b
    r=$(python bench.py --workload pmmh --chains $C --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])")
    echo "chains=$C threads=$th bpc=$b ms/iter it/s: $r" | tee -a gpurun_out/ab_chain_geom.txt
  done
done
done
