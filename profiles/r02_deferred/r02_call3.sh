#!/bin/bash
# the helper-warp pipeline: parity, then an A/B against the plain kernel and the tape without deferral
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_defer.py -x -q > gpurun_out/r02_tests3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests3.log
tail -15 gpurun_out/r02_tests3.log
for d in 1 0; do
  FFGPU_DEFER=$d timeout 300 python scripts/kbench.py --funcs f1,f2,f3,f4 --points $((1<<24)) --shapes auto --reps 5 > gpurun_out/r02_kbench_helper_defer$d.log 2>&1
  grep -E "^f[1-4]" gpurun_out/r02_kbench_helper_defer$d.log
done
timeout 300 python scripts/prof_defer.py --func f4 > gpurun_out/r02_prof_helper_plain.log 2>&1; cat gpurun_out/r02_prof_helper_plain.log
timeout 300 python scripts/prof_defer.py --func f1 >> gpurun_out/r02_prof_helper_plain.log 2>&1; tail -3 gpurun_out/r02_prof_helper_plain.log
