#!/bin/bash
# round 2, second session, call 4: bench line (N=1), reference arm, launch list, ncu --set full of f1 and f4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench rc=$?"
head -c 400 gpurun_out/r02b_bench_n1.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_bench_ref.json 2> gpurun_out/r02b_bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r02b_ncu_bench.log 2>&1
for f in f1 f4; do
  timeout 300 python scripts/prof_target.py --func $f --points $((1<<23)) > gpurun_out/r02b_prof_${f}_plain.log 2>&1 || exit 1
  cat gpurun_out/r02b_prof_${f}_plain.log
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_tape --launch-skip 2 -c 1 -f -o gpurun_out/r02b_prof_$f \
    python scripts/prof_target.py --func $f --points $((1<<23)) > gpurun_out/r02b_prof_${f}_ncu.log 2>&1
  # the report itself is too big to travel back: keep its summary and its source page
  python scripts/ncu_summary.py gpurun_out/r02b_prof_$f.ncu-rep gpurun_out/r02b_ncu_eval_tape_$f.json > /dev/null 2>&1
  ncu -i gpurun_out/r02b_prof_$f.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02b_ncu_source_$f.csv.gz
  ncu -i gpurun_out/r02b_prof_$f.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r02b_ncu_raw_$f.csv.gz
  rm -f gpurun_out/r02b_prof_$f.ncu-rep
done
