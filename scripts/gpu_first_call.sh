#!/bin/bash
# The first GPU call of a round, in one piece (run through gpurun from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/gpu_first_call.sh r02'
# 1. the whole GPU suite (the parser mirror, the add / trim / clone test and the recorded-probe test were written
#    without GPU time and run here for the first time), 2. smoke, 3. the bench line, 4. the reference arm,
# 5. the ncu launch list of the same bench command (its numbers are never bench values).
# Everything lands in gpurun_out/ with the given prefix.
set -u
tag=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_bench.log 2>&1
tail -3 gpurun_out/${tag}_tests.log gpurun_out/${tag}_smoke.log
head -c 600 gpurun_out/${tag}_bench_n1.json
