#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02b_split_launches.csv python scripts/split_latency.py > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02b_split_launches.csv")) if len(r) > 10 and r[0].isdigit()]
for r in rows[-12:]:
    print(r[4][:60], r[7], r[8], r[-1], r[-2])
PY
