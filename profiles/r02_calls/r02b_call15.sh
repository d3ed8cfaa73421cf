#!/bin/bash
# round 2, second session, call 15: whole GPU suite and the bench line with the final code
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_tests8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_tests8.log
tail -4 gpurun_out/r02b_tests8.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02b_smoke.log; tail -2 gpurun_out/r02b_smoke.log
python bench.py > gpurun_out/r02b_bench3_n1.json 2> gpurun_out/r02b_bench3_n1.err; echo "bench rc=$?"
head -c 300 gpurun_out/r02b_bench3_n1.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_bench3_ref.json 2> gpurun_out/r02b_bench3_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches3.csv \
  python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r02b_ncu_bench3.log 2>&1
