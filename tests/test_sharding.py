"""Multi-rank host logic on CPU: two gloo ranks shard a global probe batch by point index, each evaluates
its shard (with the oracle standing in for the GPU), rank 0 gathers; the result must equal the unsharded
evaluation.  This is the N>1 path of bench.py minus the device."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import firefly_b200 as ff
    from firefly_b200 import workloads as W
    from firefly_b200.shard import concat_by_rank, rank_range
    from oracle_bind import Oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = ff.primes(1)[0]
    text, vars_ = W.benchmark("f3")
    lo, hi = rank_range(total, rank, world)
    pts = ff.synthetic_points(len(vars_), hi - lo, p, seed=3, first_point=lo, total_points=total)
    out = Oracle().parse([text], vars_, p).eval(pts)
    pieces = [None] * world
    dist.all_gather_object(pieces, out)
    # max-over-ranks reduction as bench.py does for its timing
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((concat_by_rank(pieces), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_cover_the_batch(oracle):
    import firefly_b200 as ff
    from firefly_b200 import workloads as W
    total, world = 1001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    [p.start() for p in procs]
    got, tmax = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert tmax == 2.0
    p0 = ff.primes(1)[0]
    text, vars_ = W.benchmark("f3")
    full = ff.synthetic_points(len(vars_), total, p0, seed=3)
    want = oracle.parse([text], vars_, p0).eval(full)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_rank_range_partitions():
    from firefly_b200.shard import rank_range
    for total in (0, 1, 2, 7, 128, 1001):
        for world in (1, 2, 3, 4, 8):
            r = [rank_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert all(lo <= hi for lo, hi in r)
