#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for rep in 1 2; do
for v in late late_all; do
  FFGPU_LIB=$PWD/build/$v.so timeout 300 python scripts/kbench.py --funcs f1,f2,f3,f4 --points $((1<<23)) --shapes auto --reps 7 --ibp 2000 --ibp-points 65536 > gpurun_out/r02b_kbench_latecopy_${v}_$rep.log 2>&1
  echo "== $v $rep"; grep "auto" gpurun_out/r02b_kbench_latecopy_${v}_$rep.log | grep -v "^{"
done
done
for v in late late_all; do echo "== $v small"; FFGPU_LIB=$PWD/build/$v.so python scripts/split_latency.py | grep bunch; done
