#!/usr/bin/env python
"""Randomised campaign through the KERNELS (needs a GPU) -- TEST INFRASTRUCTURE.

Draws collections with the generators of scripts/fuzz_compiler.py (free grammar, shared sub-expressions, polynomial
tables with repeated denominators, IBP-like tables, long sums), compiles them with a random choice of compiler passes,
evaluates them on the GPU through ffgpu_eval at edge and random points, batch sizes 1..3000, two primes (the second
through set_prime on the same handle; now and then a prime outside FireFly's table: Barrett kernels), and compares every
value with the C oracle (oracle/ff_oracle.c).

    python scripts/fuzz_gpu.py --seconds 300 --seed 1
"""
import argparse
import os
import random
import re
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--first", type=int, default=0, help="first round (replay)")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    import firefly_b200 as ff
    from firefly_b200 import workloads as W
    from fuzz_compiler import free_grammar, long_sums, table
    from oracle_bind import Oracle
    from test_gpu_parity import _dag_expressions
    ctx = ff.Context()
    orc = Oracle()
    generic = [9223372036854742973, 4611686018427400249, 2305843009213693951, 1099511627689, 4294967311, 2097143, 65537, 997, 103]
    t0 = time.time()
    values = fails = refused = 0
    rounds = a.first
    while time.time() - t0 < a.seconds and fails < 5:
        rng = random.Random((a.seed << 20) ^ rounds)
        rounds += 1
        kind = rng.choice(["free", "dag", "table", "table", "ibp", "ibp", "long"])
        nv = rng.randint(1, 6)
        vars_ = [f"x{i}" for i in range(1, nv + 1)]
        if kind == "free":
            funs = free_grammar(rng, vars_, rng.randint(1, 40), rng.randint(1, 5))
        elif kind == "dag":
            vars_ = ["x", "y", "z", "w1", "u"][:max(nv, 2)]
            funs = _dag_expressions(rng, vars_, rng.randint(1, 30))
        elif kind == "table":
            funs = table(rng, vars_, rng.randint(2, 60))
        elif kind == "long":
            funs = long_sums(rng, vars_, rng.randint(1, 5))
        else:
            funs, vars_ = W.ibp_like_collection(rng.randint(2, 300), max(nv, 2), seed=rng.randrange(1 << 30),
                                                den_run=rng.choice([0, 0, 2, 4, 16]))
        nv = len(vars_)
        opts = dict(cse=rng.random() < 0.9, fuse=rng.random() < 0.9, cross=rng.random() < 0.9, table=rng.random() < 0.85,
                    split=rng.random() < 0.9, horner=rng.random() < 0.9, temps=rng.choice([0, 0, 0, 1, 2, 3, 8]), check_is_equal=rng.random() < 0.15)
        if opts["check_is_equal"] and any(re.search(r"\^\((?!-?\d+\))", f) for f in funs):
            opts["check_is_equal"] = False
        tag = f"seed={a.seed} round={rounds - 1} kind={kind} n={len(funs)} opts={opts}"
        if a.verbose:
            print(tag, flush=True)
        try:
            par = ff.GpuParser(ctx, funs, vars_, **opts)
        except ff.FFGpuError:
            refused += 1           # operator without operands etc. (scripts/fuzz_compiler.py checks every refusal)
            continue
        primes = [rng.choice(ff.primes()), rng.choice(ff.primes()) if rng.random() < 0.7 else rng.choice(generic)]
        try:
            for p in primes:
                par.set_prime(p)
                n = rng.choice([1, 2, 3, 7, 64, 128, 129, 1000, 3000])
                pts = ff.synthetic_points(nv, n, p, seed=rng.randrange(1 << 60))
                pts[:, 0] = 0
                if n > 3:
                    pts[:, 1] = 1
                    pts[:, 2] = p - 1
                    pts[:, 3] = pts[0, 3]
                if a.verbose:
                    print(f"   p={p} n={n} rows={par.n_table_rows} split={par.n_split_rows} instr={par.n_instr} slots={par.n_slots}", flush=True)
                try:
                    got = par.evaluate_pre(pts, thread_id=rng.randrange(64))
                except ff.FFGpuError as e:
                    print(f"FAIL {tag}: p={p} n={n}: {e}", flush=True)
                    os._exit(1)
                sel = np.unique(np.concatenate([np.arange(min(n, 6)), rng.sample(range(n), min(n, 6)), [n - 1]]))
                op = orc.parse(funs, vars_, p, check_is_equal=opts["check_is_equal"])
                want = op.eval(np.ascontiguousarray(pts[:, sel]))
                if not np.array_equal(got[:, sel], want):
                    f, j = np.argwhere(got[:, sel] != want)[0]
                    print(f"FAIL {tag}: p={p} n={n} output {f} point {sel[j]}", flush=True)
                    fails += 1
                    break
                values += want.size
                op.close()
        finally:
            par.close()
    print(f"{rounds} collections ({refused} refused as malformed), {values} function values from the GPU compared with the oracle, "
          f"{fails} failure(s), {time.time() - t0:.0f} s")
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
