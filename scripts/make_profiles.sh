#!/bin/bash
# After scripts/gpu_evidence.sh has run under gpurun: copy the bench lines and turn the ncu reports / launch lists in
# gpurun_out/ into the tracked text summaries under profiles/ (runs here, on the CPU box; ncu only reads reports).
set -e
cd "$(dirname "$0")/.."
cp gpurun_out/r1_bench_*.json profiles/ 2>/dev/null || true
for k in st_step_N2p24 st_resample_N2p24 st_step_pmmh st_resample_pmmh; do
  [ -f gpurun_out/r1_$k.ncu-rep ] || continue
  python scripts/ncu_summary.py gpurun_out/r1_$k.ncu-rep > profiles/r1_ncu_$k.txt 2>&1
  ncu -i gpurun_out/r1_$k.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,v=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','sm__cycles_elapsed.avg','sm__cycles_active.avg','sm__cycles_active.max','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__average_warp_latency_per_inst_issued.ratio','launch__grid_size','launch__block_size','launch__registers_per_thread']
print('\n== extra (raw page) ==')
for k in want:
    if k in h: i=h.index(k); print(f'{k:70s} {v[i]:>16s} {u[i]}')
print('stall reasons (warps per issue-active cycle):')
for i,k in enumerate(h):
    if 'stalled' in k and 'per_issue_active' in k and 'not_issued' not in k and float(v[i])>0.15: print('   %-24s %s' % (k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''), v[i]))
" >> profiles/r1_ncu_$k.txt
done
for k in stream_N2p24 pmmh; do
  [ -f gpurun_out/r1_ncu_launches_$k.csv ] || continue
  python scripts/launch_table.py gpurun_out/r1_ncu_launches_$k.csv > profiles/r1_ncu_launches_$k.txt
  head -c 400000 gpurun_out/r1_ncu_launches_$k.csv > profiles/r1_ncu_launches_$k.csv
done
grep -h -E "^kernel:|gpu__time_duration|dram__bytes|smsp__inst_executed.sum|issue_active" profiles/r1_ncu_st_*.txt | sort -u | head -40
