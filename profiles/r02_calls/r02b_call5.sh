#!/bin/bash
# round 2, second session, call 5 (2 GPUs): whole GPU suite (the multi-GPU context test runs), bench at N=2 (torchrun)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r02b_tests_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_tests_2gpu.log
tail -5 gpurun_out/r02b_tests_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 \
  > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err; echo "bench n2 rc=$?"
head -c 300 gpurun_out/r02b_bench_n2.json; echo
./build/h2d_peak --gpus 2 --mb 256 --reps 10 > gpurun_out/r02b_h2d_peak_n2.json 2>&1; cat gpurun_out/r02b_h2d_peak_n2.json
