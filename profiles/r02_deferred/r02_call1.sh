#!/bin/bash
# round 2, first GPU call: the GPU suite with the deferring kernel, then an A/B of the two division kernels
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests1.log
tail -5 gpurun_out/r02_tests1.log
for d in 0 1; do
  FFGPU_DEFER=$d timeout 300 python scripts/kbench.py --funcs f1,f2,f3,f4 --points $((1<<24)) --shapes auto,8x256,8x128 --reps 5 > gpurun_out/r02_kbench_defer$d.log 2>&1
  grep -E "^f[1-4]" gpurun_out/r02_kbench_defer$d.log
done
for st in 4 16; do
  FFGPU_DEFER=1 FFGPU_DEFER_STACK=$st timeout 300 python scripts/kbench.py --funcs f2,f3,f4 --points $((1<<24)) --shapes auto,8x128 --reps 5 > gpurun_out/r02_kbench_defer1_st$st.log 2>&1
  grep -E "^f[1-4]" gpurun_out/r02_kbench_defer1_st$st.log
done
