#!/bin/bash
# round 2, second session, call 23 (8 GPUs): bench at N=8 with the final code
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 \
  > gpurun_out/r02b_bench_final_n8.json 2> gpurun_out/r02b_bench_final_n8.err; echo "bench n8 rc=$?"
head -c 300 gpurun_out/r02b_bench_final_n8.json; echo
