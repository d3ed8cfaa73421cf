#!/usr/bin/env python
"""The other BASELINE.json configurations, device-resident, one JSON line each (bench.py carries configs[1]).

  --config factors   configs[2]: factors1-5, primes()[0..7], points sharded over the visible GPUs (one context,
                     contiguous point ranges per GPU, no communication); every value of a sample checked
  --config ibp       configs[3]: synthetic IBP-like collection (default 10 000 functions, 8 variables), points in
                     tiles per GPU; outputs reduced on the device to per-function (sum mod p, xor) checksums, a sample
                     of points checked value by value against the oracle
  --config sweep     configs[1] bunch sweep 1..4096 through the host API (latency view of the black-box call)
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import firefly_b200 as ff  # noqa: E402
from firefly_b200 import workloads as W  # noqa: E402
from oracle_bind import Oracle  # noqa: E402


def per_gpu_contexts(n):
    return [ff.Context([g]) for g in range(n)]


def run_factors(args):
    n_gpu = args.gpus or torch.cuda.device_count()
    primes = ff.primes()[:8]
    orc = Oracle()
    ctxs = per_gpu_contexts(n_gpu)
    n = args.points            # per GPU per (function, prime)
    res = {}
    ok = True
    t_all = 0.0
    total = 0
    for name in ["factors1", "factors2", "factors3", "factors4", "factors5"]:
        text, vars_ = W.benchmark(name)
        pars = [ff.GpuParser(c, [text], vars_) for c in ctxs]
        bufs = []
        for g in range(n_gpu):
            with torch.cuda.device(g):
                v = torch.empty((len(vars_), n), dtype=torch.int64, device=f"cuda:{g}")
                o = torch.empty((1, n), dtype=torch.int64, device=f"cuda:{g}")
                bufs.append((v, o))
        t_fn = 0.0
        for p in primes:
            for g in range(n_gpu):
                pars[g].set_prime(p)                 # prime_changed(): constants only
                with torch.cuda.device(g):
                    s = torch.cuda.current_stream(g).cuda_stream
                    v, o = bufs[g]
                    ff.lib.ffgpu_fill_points(ctxs[g].h, 0, v.data_ptr(), len(vars_), n, n, g * n, n * n_gpu, 11, p, s)
            for g in range(n_gpu):
                torch.cuda.synchronize(g)
            t0 = time.perf_counter()
            for rep in range(args.reps):
                for g in range(n_gpu):
                    with torch.cuda.device(g):
                        s = torch.cuda.current_stream(g).cuda_stream
                        v, o = bufs[g]
                        pars[g].eval_device(v.data_ptr(), n, n, o.data_ptr(), n, stream=s)
            for g in range(n_gpu):
                torch.cuda.synchronize(g)
            t_fn += (time.perf_counter() - t0) / args.reps
            # check a sample on every GPU
            idx = np.arange(0, n, max(1, n // 512))
            for g in range(n_gpu):
                host = ff.synthetic_points(len(vars_), n, p, seed=11, first_point=g * n, total_points=n * n_gpu)[:, idx]
                want = orc.parse([text], vars_, p).eval(np.ascontiguousarray(host))
                got = bufs[g][1].cpu().numpy().view(np.uint64)[:, idx]
                ok &= bool(np.array_equal(got, want))
        res[name] = {"probes_per_s": n * n_gpu * len(primes) / t_fn, "mm_eq_per_probe": pars[0].work()["mm"] + 6 * pars[0].work()["inv"]}
        t_all += t_fn
        total += n * n_gpu * len(primes)
    print(json.dumps({"config": "factors1-5 x primes()[0..7], points sharded over GPUs (BASELINE.json configs[2])",
                      "n_gpus": n_gpu, "points_per_gpu_per_function_per_prime": n, "value": total / t_all,
                      "unit": "probes/s", "per_function": res, "verified_bit_exact_vs_oracle": ok,
                      "timing": "host clock around launches on all GPUs + synchronize, mean of %d repetitions" % args.reps}))
    return ok


def run_ibp(args):
    n_gpu = args.gpus or torch.cuda.device_count()
    p = ff.primes()[0]
    funs, vars_ = W.ibp_like_collection(args.functions, 8, den_run=args.den_run)
    orc = Oracle()
    ctxs = per_gpu_contexts(n_gpu)
    pars = [ff.GpuParser(c, funs, vars_, cross=not args.no_cross) for c in ctxs]
    for q in pars:
        q.set_prime(p)
    F = pars[0].n_out
    tile = args.tile
    tiles = max(1, args.points // (tile * n_gpu))           # per GPU
    n_total = tiles * tile * n_gpu
    state = []
    for g in range(n_gpu):
        with torch.cuda.device(g):
            v = torch.empty((8, tile), dtype=torch.int64, device=f"cuda:{g}")
            o = torch.empty((F, tile), dtype=torch.int64, device=f"cuda:{g}")
            sums = torch.zeros((tiles, F, 2), dtype=torch.int64, device=f"cuda:{g}")
            state.append((v, o, sums))
    for g in range(n_gpu):
        torch.cuda.synchronize(g)
    t0 = time.perf_counter()
    for t in range(tiles):
        for g in range(n_gpu):
            with torch.cuda.device(g):
                s = torch.cuda.current_stream(g).cuda_stream
                v, o, sums = state[g]
                first = (g * tiles + t) * tile
                ff.lib.ffgpu_fill_points(ctxs[g].h, 0, v.data_ptr(), 8, tile, tile, first, n_total, 5, p, s)
                pars[g].eval_device(v.data_ptr(), tile, tile, o.data_ptr(), tile, stream=s)
                ff.lib.ffgpu_checksum(ctxs[g].h, 0, o.data_ptr(), F, tile, tile, p, sums[t].data_ptr(), s)
    for g in range(n_gpu):
        torch.cuda.synchronize(g)
    dt = time.perf_counter() - t0
    # verification: the last tile of GPU 0 -- a sample of points value by value, and its checksums recomputed on the host
    v, o, sums = state[0]
    first = (0 * tiles + tiles - 1) * tile
    idx = np.arange(0, tile, max(1, tile // 64))
    host = ff.synthetic_points(8, tile, p, seed=5, first_point=first, total_points=n_total)
    want = orc.parse(funs, vars_, p).eval(np.ascontiguousarray(host[:, idx]))
    got_full = o.cpu().numpy().view(np.uint64)
    ok = bool(np.array_equal(got_full[:, idx], want))
    sm = sums[tiles - 1].cpu().numpy().view(np.uint64)
    chk_sum = np.array([int(sum(int(x) for x in row) % p) for row in got_full[:64]], dtype=np.uint64)
    ok &= bool(np.array_equal(sm[:64, 0], chk_sum)) and bool(np.array_equal(sm[:64, 1], np.bitwise_xor.reduce(got_full[:64], axis=1)))
    w = pars[0].work()
    x = pars[0].tape_work()
    print(json.dumps({"config": "synthetic IBP-like collection (BASELINE.json configs[3])" +
                                (f", denominators shared by runs of up to {args.den_run} functions" if args.den_run else "") +
                                (", FFGPU_NO_CROSS" if args.no_cross else ""),
                      "shared_denominators": pars[0].n_shared_denominators, "n_gpus": n_gpu, "functions": F,
                      "variables": 8, "points": n_total, "tile_points": tile, "value": n_total / dt, "unit": "probes/s",
                      "function_values_per_s": n_total * F / dt, "seconds": dt,
                      "tape_instructions": pars[0].n_instr, "mm_eq_per_probe": w["mm"] + 6 * w["inv"],
                      "executed_mm_eq_per_probe": x["mm"] + 6 * x["inv"],
                      "gmm_eq_per_s_per_gpu": n_total / dt * (w["mm"] + 6 * w["inv"]) / 1e9 / n_gpu,
                      "output_bytes_per_s_per_gpu": n_total * F * 8 / dt / n_gpu,
                      "verified_bit_exact_vs_oracle": ok,
                      "timing": "host clock around all tiles on all GPUs (fill + evaluate + checksum kernels) + synchronize"}))
    return ok


def run_sweep(args):
    ctx = ff.Context()
    p = ff.primes()[0]
    out = {}
    for name in ["f1", "f2", "f3", "f4"]:
        text, vars_ = W.benchmark(name)
        par = ff.GpuParser(ctx, [text], vars_)
        par.set_prime(p)
        row = {}
        for b in [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096]:
            pts = ff.synthetic_points(par.n_vars, b, p, seed=b)
            for _ in range(20):
                par.evaluate_pre(pts)
            reps = 300
            t0 = time.perf_counter()
            for _ in range(reps):
                par.evaluate_pre(pts)
            dt = (time.perf_counter() - t0) / reps
            row[str(b)] = {"us_per_call": dt * 1e6, "probes_per_s": b / dt}
        out[name] = row
    print(json.dumps({"config": "bunch sweep 1..4096 through ffgpu_eval, pageable host buffers, one caller (BASELINE.json configs[1])",
                      "unit": "probes/s", "sweep": out}))
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["factors", "ibp", "sweep"])
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--points", type=int, default=1 << 22)
    ap.add_argument("--functions", type=int, default=10000)
    ap.add_argument("--tile", type=int, default=1 << 15)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--den-run", type=int, default=0, help="ibp: runs of up to this many functions share a denominator")
    ap.add_argument("--no-cross", action="store_true", help="ibp: no sharing across functions (A/B)")
    a = ap.parse_args()
    good = {"factors": run_factors, "ibp": run_ibp, "sweep": run_sweep}[a.config](a)
    sys.exit(0 if good else 1)
