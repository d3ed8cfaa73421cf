#!/usr/bin/env python
"""Profiling target (development tool): one function, device resident, three launches after the warm-up --
the deferring kernel, the plain kernel on the same tape, the tape compiled without deferred divisions."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import firefly_b200 as ff  # noqa: E402
from firefly_b200 import workloads as W  # noqa: E402

P0 = 9223372036854775783
ap = argparse.ArgumentParser()
ap.add_argument("--func", default="f4")
ap.add_argument("--points", type=int, default=1 << 23)
ap.add_argument("--warm", type=int, default=2)
a = ap.parse_args()
ctx = ff.Context()
text, vars_ = W.benchmark(a.func)
n = a.points
s = torch.cuda.current_stream().cuda_stream
vals = torch.empty((len(vars_), n), dtype=torch.int64, device="cuda")
out = torch.empty((1, n), dtype=torch.int64, device="cuda")
ff.lib.ffgpu_fill_points(ctx.h, 0, vals.data_ptr(), len(vars_), n, n, 0, n, 1, P0, s)
d = ff.GpuParser(ctx, [text], vars_)
nd = ff.GpuParser(ctx, [text], vars_, defer=False)
d.set_prime(P0)
nd.set_prime(P0)
runs = [(d, 1), (d, 0), (nd, 0)]
for _ in range(a.warm):
    for par, mode in runs:
        par.set_defer(mode)
        par.eval_device(vals.data_ptr(), n, n, out.data_ptr(), n, stream=s)
torch.cuda.synchronize()
for par, mode in runs:
    par.set_defer(mode)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    par.eval_device(vals.data_ptr(), n, n, out.data_ptr(), n, stream=s)
    e1.record()
    torch.cuda.synchronize()
    print(a.func, "deferred tape" if par is d else "plain tape", "mode", mode, "%.3f ms" % e0.elapsed_time(e1), flush=True)
