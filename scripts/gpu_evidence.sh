#!/bin/bash
# One GPU box, one GPU: bench lines and ncu evidence for profiles/ (run under gpurun; everything lands in gpurun_out/).
set -x
O=gpurun_out
python bench.py --steps 5 --warmup 3 > $O/r1_bench_c2_persistent.json 2> $O/ev.err
python bench.py --engine stream --N 16777216 --T 200 --steps 3 --warmup 3 --no-cpu-baseline > $O/r1_bench_stream_N2p24.json 2>> $O/ev.err
python bench.py --engine stream --N 67108864 --T 100 --steps 3 --warmup 3 --no-cpu-baseline > $O/r1_bench_stream_N2p26.json 2>> $O/ev.err
python bench.py --workload pmmh --steps 3 --warmup 3 > $O/r1_bench_pmmh_1gpu.json 2>> $O/ev.err
python bench.py --workload sharded --steps 2 --warmup 1 > $O/r1_bench_sharded_2p28_1gpu.json 2>> $O/ev.err
# launch lists (same commands, shortened)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r1_ncu_launches_stream_N2p24.csv python bench.py --engine stream --N 16777216 --T 200 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r1_ncu_launches_pmmh.csv python bench.py --workload pmmh --T 200 --steps 1 --warmup 1 > /dev/null 2>&1
# full captures of the two streaming kernels: single big filter and the PMMH batch
ncu --set full --clock-control none --import-source on -k regex:k_st_step -s 31 -c 1 -o $O/r1_st_step_N2p24 python bench.py --engine stream --N 16777216 --T 40 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_st_resample -s 31 -c 1 -o $O/r1_st_resample_N2p24 python bench.py --engine stream --N 16777216 --T 40 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_st_step -s 31 -c 1 -o $O/r1_st_step_pmmh python bench.py --workload pmmh --T 40 --steps 1 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_st_resample -s 31 -c 1 -o $O/r1_st_resample_pmmh python bench.py --workload pmmh --T 40 --steps 1 --warmup 1 > /dev/null 2>&1
ls -la $O | tail -20
