#!/bin/bash
# per-launch durations of one bench configuration: scripts/ncu_launches.sh <tag> <bench args...>
tag=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$tag.csv python bench.py "$@" --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
python scripts/launch_table.py gpurun_out/launches_$tag.csv
