#!/bin/bash
# round 2, second session, call 10 (8 GPUs): bench at N=8 (torchrun), pinned copy rates of 1..8 GPUs at once
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 \
  > gpurun_out/r02b_bench_n8.json 2> gpurun_out/r02b_bench_n8.err; echo "bench n8 rc=$?"
head -c 300 gpurun_out/r02b_bench_n8.json; echo
timeout 300 ./build/h2d_peak --gpus 8 --mb 256 --reps 10 > gpurun_out/r02b_h2d_peak_n8.json 2>&1; cat gpurun_out/r02b_h2d_peak_n8.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --no-configs \
  > gpurun_out/r02b_bench_n4.json 2> gpurun_out/r02b_bench_n4.err; echo "bench n4 rc=$?"
head -c 200 gpurun_out/r02b_bench_n4.json; echo
