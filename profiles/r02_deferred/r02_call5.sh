#!/bin/bash
set -u
mkdir -p gpurun_out
for f in f4 f1 f2; do timeout 300 python scripts/prof_defer.py --func $f >> gpurun_out/r02_prof5.log 2>&1; done; cat gpurun_out/r02_prof5.log
