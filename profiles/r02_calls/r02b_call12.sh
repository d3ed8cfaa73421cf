#!/bin/bash
# round 2, second session, call 12: long sums cut into partial-sum rows -- tests, latency of large single functions at small bunches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long_sums or shared_rows or golden or ibp or all_shapes" > gpurun_out/r02b_tests7.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/r02b_tests7.log
tail -6 gpurun_out/r02b_tests7.log
timeout 600 python - > gpurun_out/r02b_split_latency.log 2>&1 <<'PY'
import sys, time, json, random
sys.path.insert(0, "tests")
import numpy as np
import torch
import firefly_b200 as ff
ctx = ff.Context()
p = ff.primes()[0]
g = json.load(open("tests/golden/s_y_4_v.json"))
rng = random.Random(5)
def long_fn(n_terms, nv=5):
    vs = [f"x{i}" for i in range(1, nv + 1)]
    terms = ["%d*%s" % (rng.randint(1, 10**12), "*".join(f"{v}^{rng.randint(1, 30)}" for v in rng.sample(vs, 3))) for _ in range(n_terms)]
    return ["(" + "".join(rng.choice("+-") + t for t in terms)[1:] + ")/(" + "+".join(terms[: n_terms // 4]) + ")"], vs
cases = {"s_y_4_v": (g["functions"], g["vars"])}
for nt in (1000, 10000, 100000):
    cases[f"long{nt}"] = long_fn(nt)
for name, (funs, vars_) in cases.items():
    for split in (True, False):
        t0 = time.perf_counter()
        par = ff.GpuParser(ctx, funs, vars_, split=split)
        par.set_prime(p)
        tc = time.perf_counter() - t0
        w = par.tape_work()
        line = f"{name:9s} split={int(split)} compile {tc:6.2f} s, main tape {par.n_instr:7d} instr, rows {par.n_table_rows:4d} ({par.n_split_rows} partial sums), executed mm {w['mm']:8d}:"
        for b in (1, 128, 4096):
            pts = ff.synthetic_points(par.n_vars, b, p, seed=b)
            for _ in range(3):
                got = par.evaluate_pre(pts)
            reps = 20
            t0 = time.perf_counter()
            for _ in range(reps):
                par.evaluate_pre(pts)
            line += f"  bunch {b}: {(time.perf_counter() - t0) / reps * 1e6:9.1f} us"
        print(line, flush=True)
        if split:
            ref = got
        else:
            print("   identical:", bool(np.array_equal(ref, got)), flush=True)
        par.close()
PY
cat gpurun_out/r02b_split_latency.log
timeout 300 python scripts/reconstruct_times.py s_y_4_v:1:4 s_y_4_v:4:4 2>&1 | grep "aggregator=0" > gpurun_out/r02b_reconstruct_split.log; cat gpurun_out/r02b_reconstruct_split.log
