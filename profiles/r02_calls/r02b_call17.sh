#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_tape --launch-skip 16 -c 1 -f -o gpurun_out/r02b_prof_small python scripts/split_latency.py > gpurun_out/r02b_prof_small.log 2>&1
ncu -i gpurun_out/r02b_prof_small.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r02b_prof_small_raw.csv.gz
ncu -i gpurun_out/r02b_prof_small.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02b_prof_small_source.csv.gz
rm -f gpurun_out/r02b_prof_small.ncu-rep
tail -3 gpurun_out/r02b_prof_small.log
