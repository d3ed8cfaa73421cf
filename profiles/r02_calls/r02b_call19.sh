#!/bin/bash
# round 2, second session, call 19: Horner nests -- GPU fuzz, long-function latency and throughput with and without
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 400 python scripts/fuzz_gpu.py --seconds 150 --seed 31 > gpurun_out/r02b_fuzz_gpu_horner.log 2>&1; tail -2 gpurun_out/r02b_fuzz_gpu_horner.log | cut -c1-300
timeout 600 python - > gpurun_out/r02b_horner_latency.log 2>&1 <<'PY'
import sys, time, json, random, itertools
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
import torch
import firefly_b200 as ff
ctx = ff.Context()
p = ff.primes()[0]
rng = random.Random(5)
def long_fn(n_terms, nv=5):
    vs = [f"x{i}" for i in range(1, nv + 1)]
    terms = ["%d*%s" % (rng.randint(1, 10**12), "*".join(f"{v}^{rng.randint(1, 30)}" for v in rng.sample(vs, 3))) for _ in range(n_terms)]
    return ["(" + "".join(rng.choice("+-") + t for t in terms)[1:] + ")/(" + "+".join(terms[: n_terms // 4]) + ")"], vs
def dense(nv, deg):
    vs = [f"x{i}" for i in range(1, nv + 1)]
    def poly(d):
        ts = []
        for e in itertools.product(range(d + 1), repeat=nv):
            if sum(e) > d: continue
            mon = "*".join((f"{v}^{k}" if k > 1 else v) for v, k in zip(vs, e) if k > 0)
            ts.append(str(rng.randint(1, 10**9)) + ("*" + mon if mon else ""))
        rng.shuffle(ts)
        return "+".join(ts)
    return [f"({poly(deg)})/({poly(deg // 2)})"], vs
cases = {"sparse10000": long_fn(10000), "dense_4v_deg14": dense(4, 14), "dense_3v_deg30": dense(3, 30)}
s = torch.cuda.current_stream().cuda_stream
for name, (funs, vars_) in cases.items():
    ref = None
    for kw in (dict(), dict(horner=False), dict(horner=False, table=False)):
        t0 = time.perf_counter()
        par = ff.GpuParser(ctx, funs, vars_, **kw)
        par.set_prime(p)
        tc = time.perf_counter() - t0
        w = par.tape_work()
        line = f"{name:15s} {str(kw):36s} compile {tc:5.2f} s, instr {w['instr']:6d}, mm {w['mm']:6d}, rows {par.n_table_rows:3d}:"
        for b in (1, 128, 1 << 18):
            if b > 128 and kw.get("table", True) is False and w["instr"] > 100000: continue
            stride = b + (b & 1)
            vals = torch.empty((par.n_vars, stride), dtype=torch.int64, device="cuda"); out = torch.empty((1, stride), dtype=torch.int64, device="cuda")
            ff.lib.ffgpu_fill_points(ctx.h, 0, vals.data_ptr(), par.n_vars, b, stride, 0, b, 1, p, s)
            for _ in range(2): par.eval_device(vals.data_ptr(), stride, b, out.data_ptr(), stride, stream=s)
            torch.cuda.synchronize()
            reps = 10 if b <= 128 else 3
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps): par.eval_device(vals.data_ptr(), stride, b, out.data_ptr(), stride, stream=s)
            e1.record(); torch.cuda.synchronize()
            line += f"  n={b}: {e0.elapsed_time(e1) / reps * 1e3:9.1f} us"
            if b == 128:
                got = out[:, :b].cpu().numpy().copy()
        if ref is None: ref = got
        line += "  identical" if np.array_equal(ref, got) else "  DIFFERENT"
        print(line, flush=True)
        par.close()
PY
cat gpurun_out/r02b_horner_latency.log
