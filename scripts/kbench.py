#!/usr/bin/env python
"""Kernel tuning sweep (development tool): device-resident evaluation of the benchmark functions for
every (probes per thread, threads per block) shape; prints ms, probes/s and MM-eq/s per shape and checks
a sample of every result against the oracle."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import firefly_b200 as ff  # noqa: E402
from firefly_b200 import workloads as W  # noqa: E402
from oracle_bind import Oracle  # noqa: E402

P0 = 9223372036854775783


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--funcs", default="f1,f2,f3,f4")
    ap.add_argument("--points", type=int, default=1 << 23)
    ap.add_argument("--shapes", default="4x256,4x512,8x128,8x256,2x256,2x512,4x128")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--ibp", type=int, default=0, help="also run an IBP-like collection with this many functions")
    ap.add_argument("--den-run", type=int, default=0, help="IBP-like collection: runs of up to this many functions share a denominator")
    ap.add_argument("--no-cross", action="store_true", help="no sharing across functions (A/B)")
    ap.add_argument("--no-table", action="store_true", help="no shared rows (A/B)")
    ap.add_argument("--ibp-points", type=int, default=1 << 16)
    ap.add_argument("--temps", default="0", help="comma list of temp-slot budgets (0 = automatic, -1 = CSE off)")
    args = ap.parse_args()
    ctx = ff.Context()
    orc = Oracle()
    n = args.points
    stream = torch.cuda.current_stream().cuda_stream
    res = {}
    work = []
    for name in args.funcs.split(","):
        text, vars_ = W.benchmark(name)
        work.append((name, [text], vars_, n))
    if args.ibp:
        funs, vars_ = W.ibp_like_collection(args.ibp, 8, den_run=args.den_run)
        work.append((f"ibp{args.ibp}" + (f"r{args.den_run}" if args.den_run else ""), funs, vars_, min(n, args.ibp_points)))
    work = [(f"{name}/t{t}", funs, vars_, npts, int(t)) for name, funs, vars_, npts in work for t in args.temps.split(",")]
    for name, funs, vars_, npts, temps in work:
        par = ff.GpuParser(ctx, funs, vars_, cse=temps >= 0, temps=max(temps, 0), cross=not args.no_cross, table=not args.no_table)
        par.set_prime(P0)
        w = par.work()
        mm_eq = w["mm"] + 6 * w["inv"]
        tw = par.tape_work()
        print(f"{name}: {par.n_instr} instr, spill {par.stack_depth}, temps {par.n_temps}, slots {par.n_slots}, "
              f"shared rows {par.n_table_rows}, MM-eq per probe {mm_eq} (reference tokens) / {tw['mm'] + 6 * tw['inv']} (executed)", flush=True)
        vals = torch.empty((len(vars_), npts), dtype=torch.int64, device="cuda")
        out = torch.empty((par.n_out, npts), dtype=torch.int64, device="cuda")
        ff.lib.ffgpu_fill_points(ctx.h, 0, vals.data_ptr(), len(vars_), npts, npts, 0, npts, 1, P0, stream)
        idx = np.arange(0, npts, max(1, npts // 256))
        host = ff.synthetic_points(len(vars_), npts, P0, seed=1)[:, idx]
        want = orc.parse(funs, vars_, P0).eval(np.ascontiguousarray(host))
        for shp in args.shapes.split(","):
            k, th = (0, 0) if shp == "auto" else (int(x) for x in shp.split("x"))
            par.set_launch(k, th)
            out.zero_()
            try:
                for _ in range(2):
                    par.eval_device(vals.data_ptr(), npts, npts, out.data_ptr(), npts, stream=stream)
                torch.cuda.synchronize()
            except ff.FFGpuError as e:
                print(name, shp, "FAILED", e)
                continue
            best = 1e30
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                par.eval_device(vals.data_ptr(), npts, npts, out.data_ptr(), npts, stream=stream)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            ok = np.array_equal(out.cpu().numpy().view(np.uint64)[:, idx], want)
            rate = npts / (best * 1e-3)
            res[f"{name}:{shp}"] = {"ms": best, "probes_per_s": rate, "gmm_eq_per_s": rate * mm_eq / 1e9, "ok": ok}
            print(f"{name:14s} {shp:6s} {best:9.3f} ms  {rate:11.4e} probes/s  {rate*mm_eq/1e9:8.1f} GMMeq/s  ok={ok}", flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
