#!/bin/bash
# round 2, second session, call 24 (8 GPUs): the end-to-end leg at N=8 with larger and smaller calls
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for ep in 8388608 524288; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --no-configs --no-cpu-baseline --steps 10 --e2e-points $ep \
  > gpurun_out/r02b_bench_n8_e2e_$ep.json 2> gpurun_out/r02b_bench_n8_e2e_$ep.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r02b_bench_n8_e2e_$ep.json")); e=d["e2e"]
print($ep, "e2e %.4g h2d %.1f GB/s"%(e["value"], e["h2d_gb_per_s"]), e["link_peak"], "line %.4g"%d["e2e_line"]["value"])
PY
done
