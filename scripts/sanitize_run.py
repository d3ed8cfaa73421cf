#!/usr/bin/env python
"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
golden collections through ffgpu_eval (odd and even batch sizes, several launch shapes, the many-function chunked
path), the line form, the probe queue, the device-resident path with fill + checksum.  Values are checked against the
oracle so that a sanitizer run is also a parity run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import firefly_b200 as ff  # noqa: E402
from conftest import load_golden  # noqa: E402
from firefly_b200 import workloads as W  # noqa: E402
from oracle_bind import Oracle  # noqa: E402

orc = Oracle()
ctx = ff.Context()
p = ff.primes()[1]
bad = 0
for name in ("s_y_4_v", "edge_cases", "f1", "f4"):
    g = load_golden(name)
    par = ff.GpuParser(ctx, g["functions"], g["vars"])
    par.set_prime(p)
    op = orc.parse(g["functions"], g["vars"], p)
    for k, th in ((0, 0), (1, 128), (2, 256), (4, 128), (8, 256), (8, 128)):
        par.set_launch(k, th)
        for n in (1, 7, 130, 2049):
            pts = ff.synthetic_points(par.n_vars, n, p, seed=n)
            if n > 3:
                pts[:, 1] = 0
            bad += not np.array_equal(par.evaluate_pre(pts), op.eval(pts))
    par.set_launch(0, 0)
    nv = par.n_vars
    a = np.arange(1, nv + 1, dtype=np.uint64)
    s = np.arange(5, nv + 5, dtype=np.uint64)
    t = ff.synthetic_points(1, 777, p, seed=3)[0]
    pts = np.array([[(int(a[v]) * int(tj) + int(s[v])) % p for tj in t] for v in range(nv)], dtype=np.uint64)
    bad += not np.array_equal(par.evaluate_line(a, s, t), op.eval(pts))
    q = ff.ProbeQueue(par, max_batch=256)
    first = q.push_many(np.ascontiguousarray(pts.T))
    got = {}
    while len(got) < 777:
        idx, vals = q.pop(cap=300)
        for c, i in enumerate(idx):
            got[int(i)] = vals[:, c]
    want = op.eval(pts)
    bad += not all(np.array_equal(got[first + j], want[:, j]) for j in range(777))
    q.close()
    par.close()
funs, vars_ = W.ibp_like_collection(300, 8, seed=2)
par = ff.GpuParser(ctx, funs, vars_)
par.set_prime(p)
pts = ff.synthetic_points(8, 100, p, seed=9)
bad += not np.array_equal(par.evaluate_pre(pts), orc.parse(funs, vars_, p).eval(pts))
print("sanitize_run: mismatches", bad, "launches", ff.launch_count())
sys.exit(1 if bad else 0)
