#!/usr/bin/env python
"""Replay of fuzz_gpu.py rounds with variations (development tool): python scripts/fuzz_gpu_replay.py SEED ROUND [key=value ...]"""
import os, sys, random, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import numpy as np
import firefly_b200 as ff
from firefly_b200 import workloads as W
from fuzz_compiler import free_grammar, long_sums, table
from oracle_bind import Oracle
from test_gpu_parity import _dag_expressions
orc = Oracle()
generic = [9223372036854742973, 4611686018427400249, 2305843009213693951, 1099511627689, 4294967311, 2097143, 65537, 997, 103]
seed, rnd = int(sys.argv[1]), int(sys.argv[2])
over = dict(kv.split("=") for kv in sys.argv[3:])
rng = random.Random((seed << 20) ^ rnd)
kind = rng.choice(["free", "dag", "table", "table", "ibp", "ibp", "long"])
nv = rng.randint(1, 6)
vars_ = [f"x{i}" for i in range(1, nv + 1)]
if kind == "free": funs = free_grammar(rng, vars_, rng.randint(1, 40), rng.randint(1, 5))
elif kind == "dag":
    vars_ = ["x", "y", "z", "w1", "u"][:max(nv, 2)]; funs = _dag_expressions(rng, vars_, rng.randint(1, 30))
elif kind == "table": funs = table(rng, vars_, rng.randint(2, 60))
elif kind == "long": funs = long_sums(rng, vars_, rng.randint(1, 5))
else: funs, vars_ = W.ibp_like_collection(rng.randint(2, 300), max(nv, 2), seed=rng.randrange(1 << 30), den_run=rng.choice([0, 0, 2, 4, 16]))
nv = len(vars_)
opts = dict(cse=rng.random() < 0.9, fuse=rng.random() < 0.9, cross=rng.random() < 0.9, table=rng.random() < 0.85,
            split=rng.random() < 0.9, horner=rng.random() < 0.9, temps=rng.choice([0, 0, 0, 1, 2, 3, 8]), check_is_equal=rng.random() < 0.15)
if opts["check_is_equal"] and any(re.search(r"\^\((?!-?\d+\))", f) for f in funs): opts["check_is_equal"] = False
primes = [rng.choice(ff.primes()), rng.choice(ff.primes()) if rng.random() < 0.7 else rng.choice(generic)]
for k, v in over.items():
    if k in opts: opts[k] = int(v) if k == "temps" else bool(int(v))
ns = [int(x) for x in over.get("n", "1,2,3,7,64,128,129,1000").split(",")]
if "p" in over: primes = [int(over["p"])]
ctx = ff.Context()
par = ff.GpuParser(ctx, funs, vars_, **opts)
print(kind, len(funs), opts, flush=True)
for p in primes:
    par.set_prime(p)
    op = orc.parse(funs, vars_, p, check_is_equal=opts["check_is_equal"])
    for n in ns:
        pts = ff.synthetic_points(nv, n, p, seed=n)
        pts[:, 0] = 0
        if n > 3:
            pts[:, 1] = 1; pts[:, 2] = p - 1; pts[:, 3] = pts[0, 3]
        print(f"p={p} n={n} rows={par.n_table_rows} split={par.n_split_rows} instr={par.n_instr} slots={par.n_slots}", end=" ", flush=True)
        got = par.evaluate_pre(pts)
        want = op.eval(pts)
        bad = np.argwhere(got != want)
        print("ok" if not len(bad) else f"DIFFERS at {bad[:6].tolist()} ({len(bad)} values)", flush=True)
