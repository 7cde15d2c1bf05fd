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
# the headline configuration (one filter, N = 2^20) on the chain-persistent kernel: the filter's tiles over all resident blocks
for th in 128 256; do
  for b in 148 296 592 1024; do
    export BSSM_ST_THREADS=$th BSSM_ST_BPC=$b BSSM_ST_CHAIN=1
    r=$(python bench.py --engine stream --N 1048576 --T 1000 --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value']/1e9)")
    echo "C2 one filter chain kernel threads=$th bpc=$b ms G/s: $r" | tee -a gpurun_out/ab_chain_geom.txt
  done
done
export BSSM_ST_CHAIN=0; unset BSSM_ST_BPC; unset BSSM_ST_THREADS
r=$(python bench.py --engine stream --N 1048576 --T 1000 --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value']/1e9)")
echo "C2 one filter two launches per observation ms G/s: $r" | tee -a gpurun_out/ab_chain_geom.txt
