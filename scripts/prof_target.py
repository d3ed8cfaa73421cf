#!/usr/bin/env python
"""Profiling target (development tool): one benchmark function or an IBP-like collection, points resident on the
device, `--warm` untimed evaluations and one timed one (CUDA events).  Run it plain first, then under
`ncu --set full --clock-control none --import-source on -k regex:eval_tape --launch-skip <warm x kernels> -c <kernels>`."""
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
ap.add_argument("--func", default="f4", help="f1..f4, factors1..5, or ibp<N>")
ap.add_argument("--points", type=int, default=1 << 23)
ap.add_argument("--warm", type=int, default=2)
ap.add_argument("--no-table", action="store_true")
a = ap.parse_args()
ctx = ff.Context()
if a.func.startswith("ibp"):
    funs, vars_ = W.ibp_like_collection(int(a.func[3:]), 8)
else:
    text, vars_ = W.benchmark(a.func)
    funs = [text]
n = a.points
s = torch.cuda.current_stream().cuda_stream
par = ff.GpuParser(ctx, funs, vars_, table=not a.no_table)
par.set_prime(P0)
vals = torch.empty((len(vars_), n), dtype=torch.int64, device="cuda")
out = torch.empty((par.n_out, n), dtype=torch.int64, device="cuda")
ff.lib.ffgpu_fill_points(ctx.h, 0, vals.data_ptr(), len(vars_), n, n, 0, n, 1, P0, s)
for _ in range(a.warm):
    par.eval_device(vals.data_ptr(), n, n, out.data_ptr(), n, stream=s)
torch.cuda.synchronize()
l0 = ff.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
par.eval_device(vals.data_ptr(), n, n, out.data_ptr(), n, stream=s)
e1.record()
torch.cuda.synchronize()
w = par.tape_work()
print(f"{a.func} {n} points: {e0.elapsed_time(e1):.3f} ms, {ff.launch_count() - l0} kernel(s), {par.n_instr} instr, "
      f"slots {par.n_slots}, shared rows {par.n_table_rows}, executed MM-eq per probe {w['mm'] + 6 * w['inv']}", flush=True)
