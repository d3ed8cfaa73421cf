#!/bin/bash
# round 2, second session, call 3: GPU suite with shared rows; ncu --set full of the IBP-like collection (2000 functions)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_tests2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_tests2.log
tail -4 gpurun_out/r02b_tests2.log
timeout 300 python scripts/prof_target.py --func ibp2000 --points 65536 > gpurun_out/r02b_prof_ibp_plain.log 2>&1 || exit 1
cat gpurun_out/r02b_prof_ibp_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_tape --launch-skip 4 -c 2 -f -o gpurun_out/r02b_prof_ibp2000 \
  python scripts/prof_target.py --func ibp2000 --points 65536 > gpurun_out/r02b_prof_ibp_ncu.log 2>&1
tail -3 gpurun_out/r02b_prof_ibp_ncu.log
