#!/bin/bash
# round 2, second session, call 1: block sizes up to 1024 threads and whole-tile prefetch (A/B on one box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in base pf1 pf2; do
  FFGPU_LIB=$PWD/build/$v.so timeout 300 python scripts/kbench.py --funcs f1,f2,f3,f4 --points $((1<<23)) \
    --shapes 8x256,8x512,8x1024,4x1024,4x512 --reps 5 --ibp 2000 > gpurun_out/r02b_kbench_$v.log 2>&1
  tail -3 gpurun_out/r02b_kbench_$v.log | head -2
done
