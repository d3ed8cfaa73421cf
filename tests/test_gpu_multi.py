"""Several GPUs inside one context (the Reconstructor drop-in on an 8-GPU box): ffgpu_eval shards a host batch
by contiguous point ranges.  Skipped on single-GPU boxes."""
import numpy as np
import pytest

import firefly_b200 as ff
from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_in_process_multi_gpu_sharding(oracle):
    import torch
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    g = load_golden("s_y_4_v")
    p = ff.primes()[1]
    ctx = ff.Context(list(range(n_dev)))
    assert ctx.n_devices == n_dev
    par = ff.GpuParser(ctx, g["functions"], g["vars"])
    par.set_prime(p)
    op = oracle.parse(g["functions"], g["vars"], p)
    for n in (1, 3, n_dev, 1000, 65537):
        pts = ff.synthetic_points(4, n, p, seed=n)
        assert np.array_equal(par.evaluate_pre(pts), op.eval(pts)), n
    # many-function collection at few points: the tape is chunked over grid.y on every device
    from firefly_b200 import workloads as W
    funs, vars_ = W.ibp_like_collection(500, 8, seed=9)
    par2 = ff.GpuParser(ctx, funs, vars_)
    par2.set_prime(p)
    pts = ff.synthetic_points(8, 300, p, seed=2)
    assert np.array_equal(par2.evaluate_pre(pts), oracle.parse(funs, vars_, p).eval(pts))
    par.close(); par2.close(); ctx.close()
