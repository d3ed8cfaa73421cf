#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( for t in 320 160 80 40; do for m in 768 256; do FFGPU_SPLIT_TARGET=$t FFGPU_SPLIT_MIN=$m python scripts/split_latency.py; done; done
FFGPU_NO_SPLIT=1 python scripts/split_latency.py
FFGPU_NO_TABLE=1 python scripts/split_latency.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv python scripts/split_latency.py 2>/dev/null | grep eval_tape | tail -12 | cut -d, -f5,8,9,15 ) > gpurun_out/r02b_split_sweep.log 2>&1
cat gpurun_out/r02b_split_sweep.log
