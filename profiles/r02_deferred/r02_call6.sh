#!/bin/bash
set -u
mkdir -p gpurun_out
for ns in 0 2000 5000 9000 15000; do
  echo "== stagger $ns"
  for f in f4 f2 f1; do FFGPU_STAGGER_NS=$ns timeout 300 python scripts/prof_defer.py --func $f --points $((1<<24)) 2>&1 | grep "plain tape"; done
done > gpurun_out/r02_stagger.log 2>&1
cat gpurun_out/r02_stagger.log
