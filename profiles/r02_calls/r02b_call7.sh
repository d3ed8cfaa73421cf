#!/bin/bash
# round 2, second session, call 7: tape line prefetch A/B (long tapes)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in base tp1 tp2; do
  FFGPU_LIB=$PWD/build/$v.so timeout 300 python scripts/kbench.py --funcs f1,f4 --points $((1<<22)) --shapes auto --reps 5 --ibp 2000 --ibp-points 65536 > gpurun_out/r02b_kbench_tape_$v.log 2>&1
  grep -v "^{" gpurun_out/r02b_kbench_tape_$v.log | grep "auto"
done
