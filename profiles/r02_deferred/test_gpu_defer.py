"""Deferred top-level divisions on the device (tape.hpp OP_DPARK / OP_DCHAIN / OP_DUNPARK / OP_DEND): the kernel that
runs the block's inversion of one work item behind the evaluation of the next ones -- over the groups of a collection
and over several tiles of points -- must return exactly what the plain kernel (the block waits for its inversion) and
the oracle return.
Reference semantics at stake: a/0 = 0 per probe (reference: source/FFInt.cpp:146-153), evaluate_pre's value for every
function (source/include/firefly/ShuntingYardParser.hpp:365-479)."""
import numpy as np
import pytest

import firefly_b200 as ff
from conftest import load_golden
from firefly_b200 import workloads as W

pytestmark = pytest.mark.gpu
P0 = 9223372036854775783


def _collections():
    out = {}
    for name in ("f2", "f3", "f4", "factors1", "factors3"):
        text, vars_ = W.benchmark(name)
        out[name] = ([text], vars_)
    g = load_golden("s_y_4_v")
    out["s_y_4_v"] = (g["functions"], g["vars"])
    g = load_golden("edge_cases")
    out["edge_cases"] = (g["functions"], g["vars"])
    out["ibp"] = W.ibp_like_collection(120, 8, seed=9)
    out["ibp_den_runs"] = W.ibp_like_collection(90, 8, seed=3, den_run=5)
    # zero divisors on purpose: x - y vanishes wherever the test sets x = y
    out["zero_div"] = (["(x+y)/(x-y)", "1/(x-y)", "(x*y+3)/(y-x)", "y/(x+y)", "x/(x-y)^2", "(x+1)/(x-y)", "7/(y-x)",
                        "(x-y)/(x-y)", "x/(2*x-2*y)", "(y+5)/(x-y)^3", "x^2/(x*x-y*y)"], ["x", "y"])
    return out


COLL = _collections()


@pytest.mark.parametrize("name", sorted(COLL))
@pytest.mark.parametrize("k,threads", [(8, 256), (8, 128), (4, 256), (4, 128)])
def test_deferring_kernel_matches_oracle(ctx, oracle, name, k, threads):
    funs, vars_ = COLL[name]
    p = ff.primes()[1]
    par = ff.GpuParser(ctx, funs, vars_)
    plain = ff.GpuParser(ctx, funs, vars_, defer=False)
    par.set_prime(p)
    plain.set_prime(p)
    plain.set_launch(k, threads)
    par.set_launch(k, threads)
    op = oracle.parse(funs, vars_, p)
    nv = len(vars_)
    for n, max_blocks in [(2, 0), (1000, 0), (2050, 0), (4097, 1), (12346, 3), (20000, 2), (9001, 4), (70000, 1), (50001, 5)]:
        pts = ff.synthetic_points(nv, n, p, seed=n)
        pts[:, 1] = 0
        if nv >= 2:
            pts[1, 0::7] = pts[0, 0::7]            # x = y: divisors vanish in scattered lanes
            pts[:, n // 2] = pts[0, n // 2]        # ... and all variables equal in one probe
        want = op.eval(pts)
        par.set_defer(2, max_blocks)               # the deferring kernel whenever the shape has one
        assert np.array_equal(par.evaluate_pre(pts), want), (name, n, k, threads, "deferring kernel")
        par.set_defer(0)                           # the same tape on the plain kernel
        assert np.array_equal(par.evaluate_pre(pts), want), (name, n, k, threads, "plain kernel, deferred tape")
        assert np.array_equal(plain.evaluate_pre(pts), want), (name, n, k, threads, "tape without deferral")
    if name in ("f2", "f3", "f4", "ibp", "zero_div", "s_y_4_v"):
        assert par.n_deferred_divisions > 0
    assert plain.n_deferred_divisions == 0
    par.close()
    plain.close()


@pytest.mark.parametrize("name", ["f2", "f4"])
def test_deferring_kernel_full_size(ctx, oracle, name):
    """2^23 + 2 points, device resident, automatic launch: the deferring kernel (persistent blocks pipelining over
    their tiles) against the plain kernel on EVERY point, and against the oracle on a sample."""
    import torch
    text, vars_ = W.benchmark(name)
    p = P0
    n = (1 << 23) + 2
    par = ff.GpuParser(ctx, [text], vars_)
    par.set_prime(p)
    assert par.n_deferred_divisions == 1
    vals = torch.empty((5, n), dtype=torch.int64, device="cuda")
    out_a = torch.zeros((1, n), dtype=torch.int64, device="cuda")
    out_b = torch.zeros((1, n), dtype=torch.int64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    assert ff.lib.ffgpu_fill_points(ctx.h, 0, vals.data_ptr(), 5, n, n, 0, n, 5, p, s) == 0
    vals[1, 5::1000] = vals[3, 5::1000]            # f4's divisor x4 - x2 + (..)^10 stays non-zero, f2's too; zeros:
    vals[:, 12345] = 0                             # all variables zero -> divisor zero -> 0
    par.set_defer(1)
    par.eval_device(vals.data_ptr(), n, n, out_a.data_ptr(), n, stream=s)
    par.set_defer(0)
    par.eval_device(vals.data_ptr(), n, n, out_b.data_ptr(), n, stream=s)
    torch.cuda.synchronize()
    assert torch.equal(out_a, out_b)
    idx = np.unique(np.concatenate([np.arange(0, n, 8191), [12345, 5, 1005, n - 1, n - 2]]))
    host = vals[:, idx].cpu().numpy().view(np.uint64)
    want = oracle.parse([text], vars_, p).eval(np.ascontiguousarray(host))
    assert np.array_equal(out_a[:, idx].cpu().numpy().view(np.uint64), want)
    assert int(out_a[0, 12345]) == 0
    par.close()
