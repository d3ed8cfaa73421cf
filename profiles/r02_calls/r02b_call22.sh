#!/bin/bash
# round 2, second session, call 22: one function per chunk for small batches -- A/B on large functions, parity
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
cat > /tmp/exact.py <<'PY'
import sys, time, json, random, itertools, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
import torch
import firefly_b200 as ff
from oracle_bind import Oracle
ctx = ff.Context(); orc = Oracle()
p = ff.primes()[0]
rng = random.Random(5)
def long_fn(n_terms, nv=5):
    vs = [f"x{i}" for i in range(1, nv + 1)]
    terms = ["%d*%s" % (rng.randint(1, 10**12), "*".join(f"{v}^{rng.randint(1, 30)}" for v in rng.sample(vs, 3))) for _ in range(n_terms)]
    return ["(" + "".join(rng.choice("+-") + t for t in terms)[1:] + ")/(" + "+".join(terms[: n_terms // 4]) + ")"], vs
g = json.load(open("tests/golden/s_y_4_v.json"))
cases = {"s_y_4_v": (g["functions"], g["vars"]), "sparse1000": long_fn(1000), "sparse10000": long_fn(10000), "sparse100000": long_fn(100000)}
s = torch.cuda.current_stream().cuda_stream
for name, (funs, vars_) in cases.items():
    par = ff.GpuParser(ctx, funs, vars_); par.set_prime(p)
    line = f"{os.environ.get('FFGPU_NO_EXACT_CHUNKS', '-')} {name:13s} rows {par.n_table_rows} ({par.n_split_rows} partial sums):"
    for b in (1, 128, 1024):
        stride = b + (b & 1)
        vals = torch.empty((par.n_vars, stride), dtype=torch.int64, device="cuda"); out = torch.empty((par.n_out, stride), dtype=torch.int64, device="cuda")
        ff.lib.ffgpu_fill_points(ctx.h, 0, vals.data_ptr(), par.n_vars, b, stride, 0, b, 1, p, s)
        for _ in range(3): par.eval_device(vals.data_ptr(), stride, b, out.data_ptr(), stride, stream=s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): par.eval_device(vals.data_ptr(), stride, b, out.data_ptr(), stride, stream=s)
        e1.record(); torch.cuda.synchronize()
        line += f"  n={b}: {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us"
        if b == 128 and name != "sparse100000":
            host = vals[:, :b].cpu().numpy().view(np.uint64)
            ok = np.array_equal(out[:, :b].cpu().numpy().view(np.uint64), orc.parse(funs, vars_, p).eval(np.ascontiguousarray(host)))
            line += " ok" if ok else " WRONG"
    print(line, flush=True)
PY
( python /tmp/exact.py; FFGPU_NO_EXACT_CHUNKS=1 python /tmp/exact.py ) > gpurun_out/r02b_exact_chunks.log 2>&1
cat gpurun_out/r02b_exact_chunks.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long_sums or shared_rows or shared_denominators or structural or ibp or randomised" 2>&1 | tail -3
