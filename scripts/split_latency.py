import sys, time, json, random, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, "tests")
import numpy as np
import torch
import firefly_b200 as ff
ctx = ff.Context()
p = ff.primes()[0]
g = json.load(open("tests/golden/s_y_4_v.json"))
funs, vars_ = g["functions"], g["vars"]
par = ff.GpuParser(ctx, funs, vars_)
par.set_prime(p)
print(os.environ.get("FFGPU_SPLIT_TARGET"), os.environ.get("FFGPU_SPLIT_MIN"), "rows", par.n_table_rows, "partial sums", par.n_split_rows, "main", par.n_instr, [(len(t), r, o) for t, r, o in par.table_levels()])
s = torch.cuda.current_stream().cuda_stream
for b in (1, 128):
    vals = torch.empty((4, b + (b & 1)), dtype=torch.int64, device="cuda")
    out = torch.empty((par.n_out, b + (b & 1)), dtype=torch.int64, device="cuda")
    ff.lib.ffgpu_fill_points(ctx.h, 0, vals.data_ptr(), 4, b, b + (b & 1), 0, b, 1, p, s)
    for _ in range(5):
        par.eval_device(vals.data_ptr(), b + (b & 1), b, out.data_ptr(), b + (b & 1), stream=s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        par.eval_device(vals.data_ptr(), b + (b & 1), b, out.data_ptr(), b + (b & 1), stream=s)
    e1.record()
    torch.cuda.synchronize()
    print(f"  bunch {b}: {e0.elapsed_time(e1) / 50 * 1e3:8.1f} us on the device per evaluation", flush=True)
