#!/bin/bash
# the pipelined deferring kernel: parity, then an A/B against the plain kernel
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_defer.py -x -q > gpurun_out/r02_tests2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests2.log
tail -15 gpurun_out/r02_tests2.log
for d in 0 1; do
  FFGPU_DEFER=$d timeout 300 python scripts/kbench.py --funcs f1,f2,f3,f4 --points $((1<<24)) --shapes auto,8x128 --reps 5 > gpurun_out/r02_kbench_pipe_defer$d.log 2>&1
  grep -E "^f[1-4]" gpurun_out/r02_kbench_pipe_defer$d.log
done
