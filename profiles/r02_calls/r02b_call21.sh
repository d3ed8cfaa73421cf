#!/bin/bash
# round 2, second session, call 21: final bench line, reference arm, launch list, ncu --set full of f1 / f4 / the 2000-function collection
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python bench.py > gpurun_out/r02b_bench_final_n1.json 2> gpurun_out/r02b_bench_final_n1.err; echo "bench rc=$?"
head -c 250 gpurun_out/r02b_bench_final_n1.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_bench_final_ref.json 2> gpurun_out/r02b_bench_final_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_final.csv \
  python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r02b_ncu_bench_final.log 2>&1
for f in f1 f4 ibp2000; do
  pts=$((1<<23)); skip=2; cnt=1
  if [ $f = ibp2000 ]; then pts=65536; skip=4; cnt=2; fi
  timeout 300 python scripts/prof_target.py --func $f --points $pts > gpurun_out/r02b_prof_final_${f}_plain.log 2>&1 || exit 1
  cat gpurun_out/r02b_prof_final_${f}_plain.log
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_tape --launch-skip $skip -c $cnt -f -o gpurun_out/r02b_prof_final_$f \
    python scripts/prof_target.py --func $f --points $pts > gpurun_out/r02b_prof_final_${f}_ncu.log 2>&1
  python scripts/ncu_summary.py gpurun_out/r02b_prof_final_$f.ncu-rep gpurun_out/r02b_ncu_final_$f.json > /dev/null 2>&1
  rm -f gpurun_out/r02b_prof_final_$f.ncu-rep
done
