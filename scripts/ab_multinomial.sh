#!/bin/bash
# multinomial resampling on the streaming engine: positions in line vs laid out ahead on a second stream
cd /root/repo
for a in 0 1; do
  export BSSM_ST_MN_AHEAD=$a
  r1=$(python bench.py --resample-fn multinomial --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value']/1e9)")
  r2=$(python bench.py --resample-fn multinomial --engine stream --N 16777216 --T 100 --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value']/1e9)")
  echo "ahead=$a | C2 multinomial ms G/s: $r1 | N=2^24 T=100 ms G/s: $r2" | tee -a gpurun_out/ab_multinomial.txt
done
