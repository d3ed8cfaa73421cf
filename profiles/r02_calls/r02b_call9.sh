#!/bin/bash
# round 2, second session, call 9: host-computed n_base (dispatch header) A/B on one box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for rep in 1 2; do
for v in prev new; do
  FFGPU_LIB=$PWD/build/$v.so timeout 300 python scripts/kbench.py --funcs f1,f2,f3,f4 --points $((1<<23)) --shapes auto --reps 7 --ibp 2000 --ibp-points 65536 > gpurun_out/r02b_kbench_nbase_${v}_$rep.log 2>&1
  echo "== $v $rep"; grep "auto" gpurun_out/r02b_kbench_nbase_${v}_$rep.log | grep -v "^{"
done
done
