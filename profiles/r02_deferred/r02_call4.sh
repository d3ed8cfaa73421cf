#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/prof_defer.py --func f4 > gpurun_out/r02_prof4.log 2>&1; cat gpurun_out/r02_prof4.log
timeout 300 python scripts/prof_defer.py --func f1 >> gpurun_out/r02_prof4.log 2>&1; tail -3 gpurun_out/r02_prof4.log
timeout 300 python scripts/prof_defer.py --func f2 >> gpurun_out/r02_prof4.log 2>&1; tail -3 gpurun_out/r02_prof4.log
timeout 600 python -m pytest tests/test_gpu_defer.py -x -q -k "f4 or f1 or zero_div or ibp" > gpurun_out/r02_tests4.log 2>&1; tail -3 gpurun_out/r02_tests4.log
