#!/usr/bin/env python
"""Development tool: end-to-end rate of ffgpu_eval / ffgpu_eval_line (pinned host buffers) against the pipeline
chunk size."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import firefly_b200 as ff  # noqa: E402
from firefly_b200 import workloads as W  # noqa: E402

P0 = 9223372036854775783
ctx = ff.Context()
ne = 1 << 21
for name in sys.argv[1:] or ["f1", "f4"]:
    text, vars_ = W.benchmark(name)
    par = ff.GpuParser(ctx, [text], vars_)
    par.set_prime(P0)
    nv = par.n_vars
    pin = ff.lib.ffgpu_host_alloc(nv * ne * 8)
    pt = ff.lib.ffgpu_host_alloc(ne * 8)
    po = ff.lib.ffgpu_host_alloc(ne * 8)
    arr = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_uint64)), shape=(nv, ne))
    arr[:] = ff.synthetic_points(nv, ne, P0)
    tarr = np.ctypeslib.as_array(C.cast(pt, C.POINTER(C.c_uint64)), shape=(ne,))
    tarr[:] = arr[0]
    a = np.arange(1, nv + 1, dtype=np.uint64)
    s = np.arange(7, nv + 7, dtype=np.uint64)
    for chunk in (0, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20):
        par.set_chunk(chunk)
        for mode in ("eval", "line"):
            f = (lambda: par.eval_ptr(pin, ne, po)) if mode == "eval" else (lambda: par.eval_line_ptr(a, s, pt, ne, po))
            f(); f()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                f()
            dt = (time.perf_counter() - t0) / 5
            print(f"{name} chunk={chunk:8d} {mode:5s} {ne/dt:10.4e} probes/s  {ne*8*((nv if mode=='eval' else 1)+1)/dt/1e9:6.1f} GB/s over PCIe", flush=True)
