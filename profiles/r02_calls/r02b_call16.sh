#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( for v in prev late prev late; do echo "== $v"; FFGPU_LIB=$PWD/build/$v.so python scripts/split_latency.py; done
for v in prev late; do echo "== $v f4/f1 small calls"; FFGPU_LIB=$PWD/build/$v.so python - <<'PY'
import sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
import firefly_b200 as ff
from firefly_b200 import workloads as W
ctx = ff.Context(); p = ff.primes()[0]
s = torch.cuda.current_stream().cuda_stream
for name in ("f4", "f1", "ibp300"):
    if name.startswith("ibp"):
        funs, vars_ = W.ibp_like_collection(300, 8)
    else:
        t, vars_ = W.benchmark(name); funs = [t]
    par = ff.GpuParser(ctx, funs, vars_); par.set_prime(p)
    for b in (2, 128):
        vals = torch.empty((len(vars_), b), dtype=torch.int64, device="cuda"); out = torch.empty((par.n_out, b), dtype=torch.int64, device="cuda")
        ff.lib.ffgpu_fill_points(ctx.h, 0, vals.data_ptr(), len(vars_), b, b, 0, b, 1, p, s)
        for _ in range(5): par.eval_device(vals.data_ptr(), b, b, out.data_ptr(), b, stream=s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100): par.eval_device(vals.data_ptr(), b, b, out.data_ptr(), b, stream=s)
        e1.record(); torch.cuda.synchronize()
        print(f"{name} bunch {b}: {e0.elapsed_time(e1) / 100 * 1e3:7.1f} us on the device", flush=True)
PY
done ) > gpurun_out/r02b_small_prefetch.log 2>&1
cat gpurun_out/r02b_small_prefetch.log
