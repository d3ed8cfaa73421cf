#!/bin/bash
# round 2, second session, call 2: shared rows -- GPU suite, A/B on the IBP-like collection
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_tests1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_tests1.log
tail -5 gpurun_out/r02b_tests1.log
for pts in 65536 4096 128; do
timeout 300 python scripts/kbench.py --funcs f4 --points $((1<<22)) --shapes auto,8x256,4x256 --reps 5 --ibp 2000 --ibp-points $pts > gpurun_out/r02b_kbench_table1_$pts.log 2>&1
timeout 300 python scripts/kbench.py --funcs f4 --points $((1<<22)) --shapes auto,8x256,4x256 --reps 5 --ibp 2000 --ibp-points $pts --no-table > gpurun_out/r02b_kbench_table0_$pts.log 2>&1
grep ibp gpurun_out/r02b_kbench_table1_$pts.log gpurun_out/r02b_kbench_table0_$pts.log
done
timeout 600 python scripts/kbench.py --funcs f4 --points $((1<<22)) --shapes auto --reps 3 --ibp 10000 --ibp-points 32768 > gpurun_out/r02b_kbench_table1_10k.log 2>&1
timeout 600 python scripts/kbench.py --funcs f4 --points $((1<<22)) --shapes auto --reps 3 --ibp 10000 --ibp-points 32768 --no-table > gpurun_out/r02b_kbench_table0_10k.log 2>&1
grep ibp gpurun_out/r02b_kbench_table1_10k.log gpurun_out/r02b_kbench_table0_10k.log
