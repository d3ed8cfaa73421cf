#!/bin/bash
# round 2, second session, call 6: new header / divergence tests, mailbox latency of a resident kernel, IBP after the c*x fusion
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_headers.py tests/test_gpu_parity.py -m gpu -x -q -k "headers or divergence or probe_queue_header or shared_rows or fused or ibp" > gpurun_out/r02b_tests3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_tests3.log
tail -5 gpurun_out/r02b_tests3.log
timeout 120 ./build/mailbox_latency 3000 > gpurun_out/r02b_mailbox_latency.json 2> gpurun_out/r02b_mailbox_latency.err; cat gpurun_out/r02b_mailbox_latency.json gpurun_out/r02b_mailbox_latency.err
timeout 300 python scripts/kbench.py --funcs f4 --points $((1<<22)) --shapes auto --reps 5 --ibp 2000 --ibp-points 65536 > gpurun_out/r02b_kbench_fuse_cx.log 2>&1; grep ibp gpurun_out/r02b_kbench_fuse_cx.log | head -3
timeout 300 python scripts/kbench.py --funcs f4 --points $((1<<22)) --shapes auto --reps 3 --ibp 10000 --ibp-points 32768 > gpurun_out/r02b_kbench_fuse_cx_10k.log 2>&1; grep ibp gpurun_out/r02b_kbench_fuse_cx_10k.log | head -3
