#!/bin/bash
# round 2, second session, call 8: division groups -- parity tests, A/B on the IBP-like collections, regression check on f1-f4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "division_groups or shared_rows or shared_denominators or ibp or structural" > gpurun_out/r02b_tests4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_tests4.log
tail -5 gpurun_out/r02b_tests4.log
for g in 4 8 2; do
FFGPU_GROUP_SIZE=$g timeout 300 python scripts/kbench.py --funcs f4 --points $((1<<22)) --shapes auto --reps 5 --ibp 2000 --ibp-points 65536 > gpurun_out/r02b_kbench_groups_g$g.log 2>&1; grep "^ibp" gpurun_out/r02b_kbench_groups_g$g.log | head -3
done
timeout 300 python scripts/kbench.py --funcs f4 --points $((1<<22)) --shapes auto --reps 5 --ibp 2000 --ibp-points 65536 --no-groups > gpurun_out/r02b_kbench_groups_off.log 2>&1; grep "^ibp" gpurun_out/r02b_kbench_groups_off.log | head -3
timeout 300 python scripts/kbench.py --funcs f4 --points $((1<<22)) --shapes auto --reps 3 --ibp 10000 --ibp-points 32768 > gpurun_out/r02b_kbench_groups_10k.log 2>&1; grep "^ibp" gpurun_out/r02b_kbench_groups_10k.log | head -3
timeout 300 python scripts/kbench.py --funcs f4 --points $((1<<22)) --shapes auto --reps 3 --ibp 10000 --ibp-points 32768 --no-groups > gpurun_out/r02b_kbench_groups_off_10k.log 2>&1; grep "^ibp" gpurun_out/r02b_kbench_groups_off_10k.log | head -3
for v in base; do
  FFGPU_LIB=$PWD/build/$v.so timeout 300 python scripts/kbench.py --funcs f1,f2,f3,f4 --points $((1<<23)) --shapes auto --reps 7 > gpurun_out/r02b_kbench_regress_$v.log 2>&1
  grep "auto" gpurun_out/r02b_kbench_regress_$v.log | grep -v "^{"
done
timeout 300 python scripts/kbench.py --funcs f1,f2,f3,f4 --points $((1<<23)) --shapes auto --reps 7 > gpurun_out/r02b_kbench_regress_new.log 2>&1
grep "auto" gpurun_out/r02b_kbench_regress_new.log | grep -v "^{"
