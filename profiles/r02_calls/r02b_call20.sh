#!/bin/bash
# round 2, second session, call 20 (2 GPUs): whole GPU suite with the final code (multi-GPU context test runs), bench at N=2
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r02b_tests_2gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_tests_2gpu_final.log
tail -4 gpurun_out/r02b_tests_2gpu_final.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 \
  > gpurun_out/r02b_bench_final_n2.json 2> gpurun_out/r02b_bench_final_n2.err; echo "bench n2 rc=$?"
head -c 250 gpurun_out/r02b_bench_final_n2.json; echo
