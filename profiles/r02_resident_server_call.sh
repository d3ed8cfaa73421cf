#!/bin/bash
# round 2, second session, call 11: resident server for small calls -- its test first (bounded), then the suite, latencies, black-box times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "resident_server or empty_batch or concurrent_callers" > gpurun_out/r02b_tests5.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/r02b_tests5.log
tail -15 gpurun_out/r02b_tests5.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 300 python - > gpurun_out/r02b_server_latency.log 2>&1 <<'PY'
import sys, time, threading
sys.path.insert(0, "tests")
import numpy as np
import firefly_b200 as ff
from firefly_b200 import workloads as W
ctx = ff.Context()
p = ff.primes()[0]
for name in ("f4", "f1"):
    text, vars_ = W.benchmark(name)
    par = ff.GpuParser(ctx, [text], vars_)
    par.set_prime(p)
    for mode in ("server", "launch", "aggregator"):
        par.set_server(1 if mode == "server" else 0)
        par.set_coalesce(1 if mode == "aggregator" else 0)
        for b in (1, 16, 128, 512):
            pts = ff.synthetic_points(par.n_vars, b, p, seed=b)
            for _ in range(20):
                par.evaluate_pre(pts)
            t0 = time.perf_counter()
            for _ in range(300):
                par.evaluate_pre(pts)
            print(f"{name} {mode:10s} 1 thread  bunch {b:4d}: {(time.perf_counter() - t0) / 300 * 1e6:7.1f} us per call", flush=True)
        for n_thr in (4, 16):
            batches = [ff.synthetic_points(par.n_vars, 128, p, seed=1000 + t) for t in range(n_thr)]
            def worker(t):
                for _ in range(200):
                    par.evaluate_pre(batches[t], thread_id=t)
            ths = [threading.Thread(target=worker, args=(t,)) for t in range(n_thr)]
            t0 = time.perf_counter()
            [t.start() for t in ths]
            [t.join() for t in ths]
            dt = time.perf_counter() - t0
            print(f"{name} {mode:10s} {n_thr:2d} threads bunch  128: {dt / 200 * 1e6:7.1f} us per call per thread, {n_thr * 200 / dt:9.0f} calls/s", flush=True)
    print(par.server_stats())
    par.close()
PY
cat gpurun_out/r02b_server_latency.log
for m in 1 0; do FFGPU_SERVER=$m timeout 300 python scripts/reconstruct_times.py s_y_4_v:1:4 f1:64:4 factors3:16:4 2>&1 | grep "aggregator=0" | sed "s/^/server=$m /"; done > gpurun_out/r02b_reconstruct_server.log
cat gpurun_out/r02b_reconstruct_server.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_tests6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_tests6.log
tail -4 gpurun_out/r02b_tests6.log
