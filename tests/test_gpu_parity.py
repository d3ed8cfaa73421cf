"""Parity tests proper: the sm_100a path through the C ABI against the oracle, the committed golden
vectors and size-independent properties.  All bit-exact (integer arithmetic)."""
import random
import threading

import numpy as np
import pytest

import firefly_b200 as ff
from conftest import golden_names, load_golden
from firefly_b200 import workloads as W

pytestmark = pytest.mark.gpu
P0 = 9223372036854775783


# ---- device field arithmetic against Python integers -------------------------------------------------
BOUND = (1 << 63) + (1 << 32)      # the interpreter's value representation: any representative below this


def _edge_values(p, rng, n):
    edge = [0, 1, 2, 3, p - 1, p - 2, p, p + 1, (1 << 63), (1 << 63) - 1, (1 << 63) + 1, BOUND - 1, BOUND - 2,
            (1 << 63) + (1 << 31), (1 << 63) + (1 << 31) - 1, (1 << 63) + 0xFFFFFFFF, (1 << 32), (1 << 32) - 1,
            0x7FFFFFFF00000000, 0x7FFFFFFFFFFFFFFF, 0x80000000FFFFFFFF, 0x00000000FFFFFFFF, 0x7FFFFFFF80000000,
            p >> 1, (p >> 1) + 1, (1 << 62), (1 << 62) - 1]
    vals = edge + [rng.randrange(BOUND) for _ in range(n)] + [rng.randrange(p) for _ in range(n)]
    vals += [BOUND - 1 - rng.randrange(1 << 20) for _ in range(n // 4)] + [(1 << 63) - rng.randrange(1 << 20) for _ in range(n // 4)]
    # low words next to 2^32 and high words next to 2^31: x0*y0 + d*h0 passes 2^64 (the carry output of the
    # multiplier in mul_b63_w / sqr_b63_w), the cross products reach their bound
    vals += [(rng.randrange(1 << 31) << 32) | (0xFFFFFFFF - rng.randrange(1 << 12)) for _ in range(n // 4)]
    vals += [((0x7FFFFFFF - rng.randrange(1 << 8)) << 32) | (0xFFFFFFFF - rng.randrange(1 << 8)) for _ in range(n // 4)]
    vals += [(1 << 63) | (0xFFFFFFFF - rng.randrange(1 << 12)) for _ in range(n // 8)]
    return vals


@pytest.mark.parametrize("pi", [0, 1, 150, 299])
def test_field_ops_lazy_inputs(ctx, pi):
    """add/sub/mul/neg/canon accept ANY representative below 2^63 + 2^32 (bounded lazy residues) and their
    results stay below that bound (checked by chaining: op(op(a, b), b))."""
    p = ff.primes()[pi]
    rng = random.Random(pi)
    vals = _edge_values(p, rng, 2000)
    a = np.array([rng.choice(vals) for _ in range(40000)], dtype=np.uint64)
    b = np.array([rng.choice(vals) for _ in range(40000)], dtype=np.uint64)
    # all edge x edge pairs first
    ne = 27
    a[: ne * ne] = np.repeat(np.array(vals[:ne], dtype=np.uint64), ne)
    b[: ne * ne] = np.tile(np.array(vals[:ne], dtype=np.uint64), ne)
    ai, bi = [int(x) for x in a], [int(x) for x in b]
    for op, fn in [(0, lambda x, y: (x + y) % p), (1, lambda x, y: (x - y) % p), (2, lambda x, y: x * y % p),
                   (3, lambda x, y: (-x) % p), (8, lambda x, y: x % p),
                   (10, lambda x, y: ((x + y) * y) % p), (11, lambda x, y: ((x - y) * y) % p),
                   (12, lambda x, y: ((x * y) * (x * y) - y) % p), (13, lambda x, y: ((-x) * y + x) % p)]:
        got = ctx.field_ops(op, a, b, p)
        want = np.array([fn(x, y) for x, y in zip(ai, bi)], dtype=np.uint64)
        assert np.array_equal(got, want), f"op {op} prime {p}"


@pytest.mark.parametrize("pi", [0, 299])
def test_field_inverse_and_powers(ctx, pi):
    p = ff.primes()[pi]
    rng = random.Random(100 + pi)
    vals = _edge_values(p, rng, 500)
    a = np.array([rng.choice(vals) for _ in range(3000)], dtype=np.uint64)
    e = np.array([rng.choice([0, 1, 2, 3, 20, p - 1, p - 2, p - 3, (p >> 1), (p >> 1) - 1, 0xFFFFFFFE, 0x80000001,
                              rng.randrange(p), rng.randrange(1000)]) for _ in range(3000)], dtype=np.uint64)
    ai, ei = [int(x) for x in a], [int(x) for x in e]
    inv = lambda x: pow(x % p, p - 2, p)
    assert np.array_equal(ctx.field_ops(4, a, a, p), np.array([inv(x) for x in ai], dtype=np.uint64))
    assert np.array_equal(ctx.field_ops(6, a, a, p), np.array([inv(x) for x in ai], dtype=np.uint64))
    want9 = np.array([inv(x) * inv(y) % p for x, y in zip(ai, ei)], dtype=np.uint64)
    assert np.array_equal(ctx.field_ops(9, a, e, p), want9)
    from tape_interp import pow_ff, pow_int, to_neg_int
    want5 = np.array([pow_ff(x % p, y % p, p) for x, y in zip(ai, ei)], dtype=np.uint64)
    assert np.array_equal(ctx.field_ops(5, a, e, p), want5)
    want7 = np.array([pow_int(x % p, to_neg_int(y % p), p) for x, y in zip(ai, ei)], dtype=np.uint64)
    assert np.array_equal(ctx.field_ops(7, a, e, p), want7)


# ---- golden vectors ----------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names())
def test_gpu_matches_golden(ctx, name):
    g = load_golden(name)
    par = ff.GpuParser(ctx, g["functions"], g["vars"])
    for case in g["cases"]:
        par.set_prime(case["prime"])
        got = par.evaluate_pre(case["points"])
        assert np.array_equal(got, case["values"]), (name, case["prime"])
        # bunch of one, the reference's configs[0] way of calling the black box
        for j in range(case["points"].shape[1]):
            one = par.evaluate_pre(case["points"][:, j])
            assert np.array_equal(one, case["values"][:, j]), (name, j)
    par.close()


# ---- oracle on seeded inputs, every launch shape --------------------------------------------------------
@pytest.mark.parametrize("name", ["s_y_4_v", "edge_cases", "f4", "f1"])
@pytest.mark.parametrize("k,threads", [(0, 0), (1, 128), (1, 256), (1, 512), (2, 128), (2, 256), (2, 512),
                                       (4, 128), (4, 256), (4, 512), (8, 128), (8, 256), (8, 512)])
def test_gpu_matches_oracle_all_shapes(ctx, oracle, name, k, threads):
    g = load_golden(name)
    p = ff.primes()[2]
    par = ff.GpuParser(ctx, g["functions"], g["vars"])
    par.set_prime(p)
    par.set_launch(k, threads)
    op = oracle.parse(g["functions"], g["vars"], p)
    for n in (1, 2, 3, 31, 257, 1000, 2049):
        pts = ff.synthetic_points(par.n_vars, n, p, seed=n)
        if n >= 3:
            pts[:, 1] = 0                      # zero divisors / zero bases inside a batch of good lanes
            pts[0, 2] = p - 1
        assert np.array_equal(par.evaluate_pre(pts), op.eval(pts)), (name, n, k, threads)
    par.close()


def test_gpu_random_expressions_vs_oracle(ctx, oracle):
    rng = random.Random(21)
    vars_ = ["a", "b", "c1", "d"]

    def atom():
        r = rng.random()
        if r < 0.5:
            return rng.choice(vars_)
        if r < 0.85:
            return str(rng.randint(0, 10 ** rng.randint(1, 17)))
        return str(rng.randint(10 ** 18, 10 ** 25))

    def expr(d):
        if d == 0 or rng.random() < 0.2:
            return atom()
        r = rng.random()
        if r < 0.5:
            return f"{expr(d-1)}{rng.choice('+-*/')}{expr(d-1)}"
        if r < 0.6:
            return f"({expr(d-1)})"
        if r < 0.72:
            return f"({expr(d-1)})^{rng.randint(0, 40)}"
        if r < 0.8:
            return f"{atom()}^(-{rng.randint(1, 9)})"
        if r < 0.85:
            return f"(-{atom()})^{rng.randint(1, 5)}"
        if r < 0.9:
            return f"(-({expr(d-1)}))"
        if r < 0.93:
            return f"({expr(d-1)})^({rng.choice(vars_)})"
        if r < 0.96:
            return f"({expr(d-1)})^(-{rng.choice(vars_)})"
        return f"(-{rng.choice(vars_)}^{rng.randint(2, 7)}+{expr(d-1)})"

    funs = [expr(5) for _ in range(400)]
    for p in (P0, ff.primes()[123]):
        op = oracle.parse(funs, vars_, p)
        par = ff.GpuParser(ctx, funs, vars_)
        par.set_prime(p)
        pts = ff.synthetic_points(4, 700, p, seed=33)
        pts[:, 0] = 0
        pts[:, 1] = np.array([1, p - 1, 2, 0], dtype=np.uint64)
        pts[1, 5:40] = np.arange(35, dtype=np.uint64)             # small lane exponents
        assert np.array_equal(par.evaluate_pre(pts), op.eval(pts))
        par.close()


def _dag_expressions(rng, vars_, n):
    """expressions that repeat sub-expressions on purpose (numerator/denominator sharing sums, repeated
    powers and products): what the compiler's sharing pass (temp slots, co-evaluated sums) feeds on"""
    def power():
        return f"{rng.choice(vars_)}^{rng.randint(2, 30)}"

    def product():
        return "*".join(rng.sample(vars_, rng.randint(2, len(vars_))))

    out = []
    for _ in range(n):
        pool = [power() for _ in range(rng.randint(2, 6))] + [product() for _ in range(rng.randint(1, 3))]
        pool += [f"({rng.choice(pool)}+{rng.randint(1, 99)})^{rng.randint(2, 9)}" for _ in range(2)]

        def ssum(k):
            terms = [rng.choice(pool) for _ in range(k)]
            return terms[0] + "".join(rng.choice("+-") + t for t in terms[1:])

        kind = rng.randrange(5)
        if kind == 0:
            e = f"({ssum(rng.randint(2, 8))})/({ssum(rng.randint(2, 8))})"
        elif kind == 1:
            e = f"({ssum(4)})*({ssum(4)})-({ssum(3)})*({rng.choice(pool)})"
        elif kind == 2:
            s1 = ssum(rng.randint(3, 6))
            e = f"({s1})*({s1})+({s1})/({ssum(3)})"
        elif kind == 3:
            q = rng.choice(pool)
            e = f"({q})+({q})^2+({q})^3-({q})*({ssum(2)})"
        else:
            e = f"(({ssum(5)})-({ssum(5)}))/(({ssum(5)})+({ssum(5)})*({rng.choice(pool)}))"
        out.append(e)
    return out


@pytest.mark.parametrize("temps", [0, 1, 2, 3, 8])
def test_shared_subexpressions_vs_oracle(ctx, oracle, temps):
    rng = random.Random(5150 + temps)
    vars_ = ["x", "y", "z", "w1", "u"]
    funs = _dag_expressions(rng, vars_, 300)
    for name in ("f1", "f2", "factors3", "factors5"):
        text, v = W.benchmark(name)
        if len(v) <= 5:
            funs.append(text.replace("x1", "x").replace("x2", "y").replace("x3", "z").replace("x4", "w1").replace("x5", "u"))
    p = ff.primes()[7]
    op = oracle.parse(funs, vars_, p)
    pts = ff.synthetic_points(5, 515, p, seed=temps)
    pts[:, 3] = 0
    pts[:, 4] = np.array([1, 1, p - 1, 2, 0], dtype=np.uint64)
    want = op.eval(pts)
    par = ff.GpuParser(ctx, funs, vars_, temps=temps)
    par.set_prime(p)
    assert par.n_temps > 0 and par.cse_fallbacks == 0       # the sharing pass found something and held up
    assert np.array_equal(par.evaluate_pre(pts), want)
    plain = ff.GpuParser(ctx, funs, vars_, cse=False)
    plain.set_prime(p)
    assert plain.n_temps == 0 and plain.n_instr > par.n_instr
    assert np.array_equal(plain.evaluate_pre(pts), want)
    unfused = ff.GpuParser(ctx, funs, vars_, temps=temps, fuse=False)      # same tape, one dispatch per operation
    unfused.set_prime(p)
    assert unfused.n_instr > par.n_instr
    assert np.array_equal(unfused.evaluate_pre(pts), want)
    par.close()
    plain.close()
    unfused.close()


@pytest.mark.parametrize("name", ["f1", "f3", "factors4", "s_y_4_v"])
def test_eval_line_matches_expanded_points(ctx, oracle, name):
    """ffgpu_eval_line builds x_v = a_v*t + s_v on the GPU (Reconstructor::queue_probes, :2946-2960): same values
    as evaluating the host-expanded points, and as the oracle."""
    g = load_golden(name)
    rng = random.Random(hash(name) & 0xffff)
    for p in (P0, ff.primes()[200]):
        par = ff.GpuParser(ctx, g["functions"], g["vars"])
        par.set_prime(p)
        par.set_chunk(1 << 12)                  # several pipelined chunks
        nv = par.n_vars
        op = oracle.parse(g["functions"], g["vars"], p)
        for n in (1, 2, 7, 4096, 10001):
            a = np.array([1] + [rng.randrange(p) for _ in range(nv - 1)], dtype=np.uint64)
            s = np.array([rng.randrange(p) for _ in range(nv)], dtype=np.uint64)
            if n == 7:                          # factor-scan shape: a in {0, 1}, fixed values elsewhere
                a = np.array([rng.randrange(2) for _ in range(nv)], dtype=np.uint64)
                s = np.where(a == 1, 0, s).astype(np.uint64)
            t = ff.synthetic_points(1, n, p, seed=n)[0]
            t[0] = 0
            if n > 2:
                t[1], t[2] = p - 1, 1
            pts = np.array([[(int(a[v]) * int(tj) + int(s[v])) % p for tj in t] for v in range(nv)], dtype=np.uint64)
            want = op.eval(pts)
            assert np.array_equal(par.evaluate_pre(pts), want)
            assert np.array_equal(par.evaluate_line(a, s, t), want), (name, p, n)
        par.close()


def test_eval_line_rejects_bad_parameters(ctx):
    par = ff.GpuParser(ctx, ["x+y"], ["x", "y"])
    par.set_prime(P0)
    with pytest.raises(ff.FFGpuError, match="below the prime"):
        par.evaluate_line([1, P0], [0, 0], [1, 2, 3])
    with pytest.raises(ValueError):
        par.evaluate_line([1], [0, 0], [1, 2, 3])


def test_check_is_equal_and_prime_change(ctx, oracle):
    funs = ["x+y", "y+x", "x*y/3", "(x+y)", "x*y/3*1", "x-y/7"]
    par = ff.GpuParser(ctx, funs, ["x", "y"], check_is_equal=True)
    assert par.n_functions == 3 and par.n_out == 6
    for p in ff.primes()[:5]:
        par.set_prime(p)                       # prime_changed(): constants only
        pts = ff.synthetic_points(2, 100, p)
        want = oracle.parse(funs, ["x", "y"], p, check_is_equal=True).eval(pts)
        assert np.array_equal(par.evaluate_pre(pts), want)


@pytest.mark.parametrize("coalesce,cap", [(1, 0), (1, 130), (0, 0)])
def test_concurrent_callers(ctx, oracle, coalesce, cap):
    """Reconstructor calls the black box from thr_n threads at once (Reconstructor.hpp:2981), each with a bunch of at
    most 128 points (utils.cpp:305-330).  With the aggregator on, calls that arrive while a launch is in flight share
    the next one: every caller must still get exactly its own values -- also when the shared batch is too small for
    all of them (cap 130: two bunches of 77 do not fit, the second waits for the next batch)."""
    g = load_golden("s_y_4_v")
    par = ff.GpuParser(ctx, g["functions"], g["vars"])
    par.set_prime(P0)
    par.set_coalesce(coalesce, cap)
    op = oracle.parse(g["functions"], g["vars"], P0)
    errs = []
    n_threads, n_iter = 32, 25
    l0 = ff.launch_count()

    def worker(tid):
        try:
            for it in range(n_iter):
                n = [1, 4, 128, 77, 2, 33][(it + tid) % 6]
                pts = ff.synthetic_points(4, n, P0, seed=1000 * tid + it)
                if n > 2:
                    pts[:, 1] = 0                  # a zero divisor in somebody's slice must stay in that slice
                got = par.evaluate_pre(pts, thread_id=tid)
                if not np.array_equal(got, op.eval(pts)):
                    errs.append((tid, it))
        except Exception as e:  # pragma: no cover
            errs.append((tid, repr(e)))

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    st = par.coalesce_stats()
    kernels = 1 + len(par.table_levels())        # a collection with shared rows runs their tapes first
    launches = (ff.launch_count() - l0) // kernels
    if coalesce:
        assert st["calls"] == n_threads * n_iter and st["launches"] == launches
        assert launches < n_threads * n_iter, "32 concurrent callers never shared a launch"
    else:
        assert st["calls"] == 0 and launches == n_threads * n_iter
    par.close()


def test_aggregator_follows_a_rebuilt_collection(ctx, oracle):
    """add_otf / trim change the number of outputs: the aggregator's staging is laid out again for the new shape"""
    funs, vars_ = ["x+y", "x*y/(x-y)"], ["x", "y"]
    par = ff.GpuParser(ctx, funs, vars_)
    par.set_prime(P0)
    pts = ff.synthetic_points(2, 17, P0, seed=3)
    assert np.array_equal(par.evaluate_pre(pts), oracle.parse(funs, vars_, P0).eval(pts))
    par.add_otf("x^3-y")
    par.set_prime(P0)
    funs2 = funs + ["x^3-y"]
    for n in (17, 200):
        pts = ff.synthetic_points(2, n, P0, seed=n)
        assert np.array_equal(par.evaluate_pre(pts), oracle.parse(funs2, vars_, P0).eval(pts))
    par.close()


# ---- full-size properties ----------------------------------------------------------------------------
def test_large_batch_sample_and_properties(ctx, oracle):
    """2^22 points of f4 through the chunked host path: a strided sample against the oracle, and the
    identity N(x)/D(x) * D(x) == N(x) for every point (size-independent, catches any lane mix-up)."""
    text, vars_ = W.benchmark("f4")
    p = P0
    n = 1 << 22
    par = ff.GpuParser(ctx, [text, "(1+x1+x2+x3+x4+x5)^20-1", "x4-x2+(x1*x2*x3*x4*x5)^10"], vars_)
    par.set_prime(p)
    par.set_chunk(1 << 19)
    pts = ff.synthetic_points(5, n, p, seed=77)
    out = par.evaluate_pre(pts)
    idx = np.arange(0, n, 4099)
    op = oracle.parse([text, "(1+x1+x2+x3+x4+x5)^20-1", "x4-x2+(x1*x2*x3*x4*x5)^10"], vars_, p)
    assert np.array_equal(out[:, idx], op.eval(np.ascontiguousarray(pts[:, idx])))
    # q * D == N for all points where D != 0 (object arrays: exact integers)
    q, N, D = (out[i].astype(object) for i in range(3))
    assert np.all((q * D - N) % p == 0)
    par.close()


def test_device_resident_path_and_checksum(ctx, oracle):
    import torch
    text, vars_ = W.benchmark("f3")
    p = ff.primes()[4]
    par = ff.GpuParser(ctx, [text], vars_)
    par.set_prime(p)
    for n in (1 << 16, (1 << 16) + 1, 12345):
        stride = n + (n & 1)
        vals = torch.empty((5, stride), dtype=torch.int64, device="cuda")
        out = torch.zeros((1, stride), dtype=torch.int64, device="cuda")
        s = torch.cuda.current_stream().cuda_stream
        assert ff.lib.ffgpu_fill_points(ctx.h, 0, vals.data_ptr(), 5, n, stride, 0, n, 99, p, s) == 0
        par.eval_device(vals.data_ptr(), stride, n, out.data_ptr(), stride, stream=s)
        sums = torch.zeros(2, dtype=torch.int64, device="cuda")
        assert ff.lib.ffgpu_checksum(ctx.h, 0, out.data_ptr(), 1, n, stride, p, sums.data_ptr(), s) == 0
        torch.cuda.synchronize()
        host_pts = ff.synthetic_points(5, n, p, seed=99)
        assert np.array_equal(vals[:, :n].cpu().numpy().view(np.uint64), host_pts)
        want = oracle.parse([text], vars_, p).eval(host_pts)
        got = out[:, :n].cpu().numpy().view(np.uint64)
        assert np.array_equal(got, want)
        sm = sums.cpu().numpy().view(np.uint64)
        assert int(sm[0]) == sum(int(x) for x in want[0]) % p
        assert int(sm[1]) == int(np.bitwise_xor.reduce(want[0]))


def test_ibp_like_collection(ctx, oracle):
    funs, vars_ = W.ibp_like_collection(600, 8, seed=5)
    p = P0
    par = ff.GpuParser(ctx, funs, vars_)
    par.set_prime(p)
    op = oracle.parse(funs, vars_, p)
    pts = ff.synthetic_points(8, 513, p, seed=4)
    pts[:, 7] = 0
    assert np.array_equal(par.evaluate_pre(pts), op.eval(pts))


def test_empty_batch_and_missing_prime(ctx):
    par = ff.GpuParser(ctx, ["x"], ["x"])
    with pytest.raises(ff.FFGpuError, match="set_prime"):
        par.evaluate_pre(np.array([[1]], dtype=np.uint64))
    par.set_prime(P0)
    out = par.evaluate_pre(np.zeros((1, 0), dtype=np.uint64))
    assert out.shape == (1, 0)
    n0 = ff.launch_count()
    par.evaluate_pre(np.array([[5]], dtype=np.uint64))
    assert ff.launch_count() == n0 + 1


# ---- the probe queue (requested_probes / computed_probes with one GPU-feeding consumer) ---------------------
def test_probe_queue_matches_oracle_by_index(ctx, oracle):
    g = load_golden("s_y_4_v")
    p = ff.primes()[5]
    par = ff.GpuParser(ctx, g["functions"], g["vars"])
    par.set_prime(p)
    op = oracle.parse(g["functions"], g["vars"], p)
    q = ff.ProbeQueue(par, max_batch=512)
    n_threads, per_thread = 6, 300
    pts = ff.synthetic_points(4, n_threads * per_thread + 1000 + 777, p, seed=8)
    pts[:, 5] = 0
    owner = {}                                  # index -> column of pts
    lock = threading.Lock()

    def producer(tid):
        for i in range(per_thread):
            col = tid * per_thread + i
            idx = q.push(pts[:, col])
            with lock:
                owner[idx] = col

    ts = [threading.Thread(target=producer, args=(t,)) for t in range(n_threads)]
    [t.start() for t in ts]
    # a block of points at once (point-major, the reference's queue layout) while the producers run
    base = n_threads * per_thread
    first = q.push_many(np.ascontiguousarray(pts[:, base:base + 1000].T))
    with lock:
        for i in range(1000):
            owner[first + i] = base + i
    [t.join() for t in ts]
    # and a line batch: only t travels
    rng = random.Random(3)
    a = np.array([1] + [rng.randrange(p) for _ in range(3)], dtype=np.uint64)
    s = np.array([rng.randrange(p) for _ in range(4)], dtype=np.uint64)
    tvals = pts[0, base + 1000:base + 1777].copy()
    first_line = q.push_line(a, s, tvals)
    line_pts = np.array([[(int(a[v]) * int(tj) + int(s[v])) % p for tj in tvals] for v in range(4)], dtype=np.uint64)
    total = n_threads * per_thread + 1000 + 777
    got = {}
    while len(got) < total:
        idx, vals = q.pop(cap=257)
        assert len(idx) > 0, "queue ran dry before all probes came back"
        for k, i in enumerate(idx):
            assert int(i) not in got
            got[int(i)] = vals[:, k]
    assert q.pending == 0 and len(q.pop(wait=False)[0]) == 0
    assert sorted(got) == list(range(total))    # every index exactly once
    want = op.eval(pts[:, :base + 1000])
    for i, col in owner.items():
        assert np.array_equal(got[i], want[:, col]), (i, col)
    want_line = op.eval(line_pts)
    for k in range(777):
        assert np.array_equal(got[first_line + k], want_line[:, k])
    st = q.stats()
    assert st["probes"] == total and st["largest_batch"] <= 1000 and st["batches"] < total   # batching happened
    q.close()
    par.close()


def test_probe_queue_reports_evaluation_errors(ctx):
    par = ff.GpuParser(ctx, ["x+y"], ["x", "y"])      # no prime set: every evaluation fails
    q = ff.ProbeQueue(par)
    q.push([1, 2])
    with pytest.raises(ff.FFGpuError, match="set_prime"):
        q.pop()
    q.close()


def test_offline_probe_service(ctx, oracle, tmp_path):
    """firefly_b200.offline: requested_probes.gz -> probes/<k>.gz for the stock `firefly` executable"""
    import gzip
    from firefly_b200 import offline
    g = load_golden("s_y_4_v")
    fn = tmp_path / "funs.m"
    fn.write_text(";\n".join(g["functions"]) + ";\n")
    k = 3
    p = ff.primes()[k]
    pts = ff.synthetic_points(4, 300, p, seed=41)
    pts[:, 0] = 0
    with gzip.open(tmp_path / "requested_probes.gz", "wt") as f:
        for j in range(300):
            f.write(f"{1 + j % 3} 1 2 | {int(pts[0, j])} | " + " ".join(str(int(x)) for x in pts[:, j]) + "\n")
    n = offline.serve(str(fn), g["vars"], k, str(tmp_path / "requested_probes.gz"), str(tmp_path / "probes" / f"{k}.gz"))
    assert n == 300
    orders, ts, vals = offline.read_probes(str(tmp_path / "probes" / f"{k}.gz"))
    assert orders[4] == "2 1 2" and ts == [int(x) for x in pts[0]]
    assert np.array_equal(vals, oracle.parse(g["functions"], g["vars"], p).eval(pts))


def test_mpi_wire_format(ctx, oracle):
    """[index, x_1..x_n]*k in, [index, f_1..f_m]*k out -- what an MPIWorker exchanges with the master"""
    g = load_golden("s_y_4_v")
    p = ff.primes()[2]
    par = ff.GpuParser(ctx, g["functions"], g["vars"])
    par.set_prime(p)
    k = 333
    pts = ff.synthetic_points(4, k, p, seed=17)
    idx = np.arange(1000, 1000 + k, dtype=np.uint64)[::-1].copy()
    msg = np.concatenate([idx.reshape(-1, 1), pts.T], axis=1)
    res = par.evaluate_wire(msg).reshape(k, par.n_out + 1)
    assert np.array_equal(res[:, 0], idx)
    assert np.array_equal(res[:, 1:].T, oracle.parse(g["functions"], g["vars"], p).eval(pts))
    with pytest.raises(ValueError, match="Corrupted"):
        par.evaluate_wire(np.zeros(7, dtype=np.uint64))


def test_scanned_factors_on_device(ctx, oracle):
    """ffgpu_set_factors: out_i = f_i / factor_i evaluated inside the tape (Reconstructor.hpp:3449-3457)"""
    g = load_golden("s_y_4_v")
    p = ff.primes()[6]
    par = ff.GpuParser(ctx, g["functions"], g["vars"])
    par.set_prime(p)
    pts = ff.synthetic_points(4, 1500, p, seed=23)
    pts[:, 3] = np.array([1, p - 1, 0, 0], dtype=np.uint64)          # x1 + x2 = 0: factor of output 2 vanishes
    base = oracle.parse(g["functions"], g["vars"], p).eval(pts)
    assert np.array_equal(par.evaluate_pre(pts), base)
    facs = {2: "(x1+x2)*(x3^2+1)", 5: "x4/(x1-x2)", 18: "3"}
    fvars = ["x1", "x2", "x3", "x4"]
    par.set_factors(facs, fvars)
    with pytest.raises(ff.FFGpuError, match="set_prime"):
        par.evaluate_pre(pts)
    par.set_prime(p)
    got = par.evaluate_pre(pts)
    fv = oracle.parse(list(facs.values()), fvars, p).eval(pts)
    want = base.copy()
    for row, i in enumerate(facs):
        want[i] = [int(b) * pow(int(f), p - 2, p) % p for b, f in zip(base[i], fv[row])]
    assert np.array_equal(got, want)
    par.set_factors({}, fvars)
    par.set_prime(p)
    assert np.array_equal(par.evaluate_pre(pts), base)
    par.close()


# ---- small-call transfer modes (capi.cu, FFGPU_SMALL_ZERO_COPY) ----------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["0", "1", "2"])
def test_small_call_transfer_modes(mode):
    """Bunches of 1..129 points take the latency path; copying both ways, letting the kernel write the results into
    mapped pinned memory (default) and letting it read the coordinates from there as well give the golden values."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, 'tests')\n"
        "import firefly_b200 as ff\n"
        "from conftest import load_golden\n"
        "ctx = ff.Context()\n"
        "for name in ('s_y_4_v', 'edge_cases'):\n"
        "    g = load_golden(name)\n"
        "    par = ff.GpuParser(ctx, g['functions'], g['vars'])\n"
        "    for case in g['cases']:\n"
        "        par.set_prime(case['prime'])\n"
        "        for n in (1, 2, 3, case['points'].shape[1]):\n"
        "            pts = np.ascontiguousarray(case['points'][:, :n])\n"
        "            assert np.array_equal(par.evaluate_pre(pts), case['values'][:, :n]), (name, n)\n"
        "        rep = np.ascontiguousarray(np.tile(case['points'], (1, 17))[:, :129])\n"
        "        assert np.array_equal(par.evaluate_pre(rep), np.tile(case['values'], (1, 17))[:, :129]), name\n"
        "print('ok')\n")
    env = dict(os.environ, FFGPU_SMALL_ZERO_COPY=mode)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr


@pytest.mark.gpu
def test_structural_duplicates_on_device(ctx, oracle):
    """functions of equal structure are evaluated once and stored to every output (tests/test_compiler.py has the
    host-side twin of this test)"""
    vars_ = ["x", "y", "z"]
    base = ["(x+y)*z/(x-3)", "x^2+y", "z*(y+x)/(x-3)", "(x+y)*z/(x-4)", "y+x^2", "(x+y)*z/(x-3)", "(x-y)*z/(x-3)"]
    funs = base * 40                                   # 280 outputs, 4 distinct structures: also crosses tape chunks
    p = ff.primes()[2]
    par = ff.GpuParser(ctx, funs, vars_)
    par.set_prime(p)
    assert par.n_structural_duplicates == len(funs) - 4
    for n in (1, 7, 4096, 70001):
        pts = ff.synthetic_points(3, n, p, seed=n)
        pts[0, 0] = 3
        sel = np.unique(np.concatenate([np.arange(min(n, 64)), np.arange(0, n, max(1, n // 97)), [n - 1]]))
        got = par.evaluate_pre(pts)[:, sel]
        assert np.array_equal(got, oracle.parse(funs, vars_, p).eval(np.ascontiguousarray(pts[:, sel]))), n
    par.close()


@pytest.mark.gpu
@pytest.mark.parametrize("den_run", [3, 12])
def test_shared_denominators_on_device(ctx, oracle, den_run):
    """runs of functions over one denominator invert it once (temp slot 0 carries 1/D from function to function);
    small batches cut the tape into many chunks (grid.y), which must never start inside a run"""
    funs, vars_ = W.ibp_like_collection(700, 8, seed=77 + den_run, den_run=den_run)
    funs = funs[:350] + funs[100:130] + funs[350:]            # and some structural duplicates in between
    if den_run == 12:
        random.Random(5).shuffle(funs)                        # equal denominators need not be neighbours
    p = ff.primes()[5]
    par = ff.GpuParser(ctx, funs, vars_)
    par.set_prime(p)
    assert par.n_shared_denominators > 100 and par.n_structural_duplicates >= 30 and par.cse_fallbacks == 0
    orc = oracle.parse(funs, vars_, p)
    for n in (1, 2, 5, 33, 300, 5000, 40001):
        pts = ff.synthetic_points(8, n, p, seed=n)
        sel = np.unique(np.concatenate([np.arange(min(n, 8)), np.arange(0, n, max(1, n // 13)), [n - 1]]))
        got = par.evaluate_pre(pts)[:, sel]
        assert np.array_equal(got, orc.eval(np.ascontiguousarray(pts[:, sel]))), n
    # every launch shape, chunked and not
    pts = ff.synthetic_points(8, 3000, p, seed=1)
    want = orc.eval(np.ascontiguousarray(pts[:, :40]))
    for k, th in [(1, 128), (2, 256), (4, 512), (8, 128), (8, 256)]:
        par.set_launch(k, th)
        assert np.array_equal(par.evaluate_pre(pts)[:, :40], want), (k, th)
    par.close()


@pytest.mark.gpu
@pytest.mark.parametrize("generic", [False, True])
def test_shared_rows_on_device(ctx, oracle, generic):
    """sub-expressions of several functions (powers of the variables, common factors) are evaluated once per probe
    into scratch rows by a pre-pass kernel and read by the functions' tape as extra variables: every path that
    launches (small call, chunked host call, caller's device memory, ranges of a big batch, every launch shape),
    against the oracle and against the same collection compiled without the pass"""
    import torch
    funs, vars_ = W.ibp_like_collection(500, 8, seed=2026, den_run=3)
    funs = funs[:200] + funs[20:30] + funs[200:]
    p = 4611686018427400249 if generic else ff.primes()[7]     # a 62-bit prime: Barrett kernels
    par = ff.GpuParser(ctx, funs, vars_)
    off = ff.GpuParser(ctx, funs, vars_, table=False)
    par.set_prime(p)
    off.set_prime(p)
    assert par.n_table_rows > 40 and off.n_table_rows == 0 and par.cse_fallbacks == 0
    wp, wo = par.tape_work(), off.tape_work()
    assert wp["mm"] + 6 * wp["inv"] < 0.8 * (wo["mm"] + 6 * wo["inv"])
    orc = oracle.parse(funs, vars_, p)
    for n in (1, 2, 7, 128, 129, 4000, 30001):
        pts = ff.synthetic_points(8, n, p, seed=n)
        if n > 2:
            pts[:, 1] = 0
            pts[:, 2] = p - 1
        sel = np.unique(np.concatenate([np.arange(min(n, 8)), np.arange(0, n, max(1, n // 11)), [n - 1]]))
        got = par.evaluate_pre(pts)
        assert np.array_equal(got[:, sel], orc.eval(np.ascontiguousarray(pts[:, sel]))), n
        assert np.array_equal(got, off.evaluate_pre(pts)), n
    # a batch whose rows exceed the scratch budget is evaluated range by range (caller's device memory, odd size)
    n = 200001 if not generic else 60001
    stride = n + 1
    s = torch.cuda.current_stream().cuda_stream
    vals = torch.empty((8, stride), dtype=torch.int64, device="cuda")
    assert ff.lib.ffgpu_fill_points(ctx.h, 0, vals.data_ptr(), 8, n, stride, 0, n, 5, p, s) == 0
    out = torch.zeros((par.n_out, stride), dtype=torch.int64, device="cuda")
    out2 = torch.zeros((par.n_out, stride), dtype=torch.int64, device="cuda")
    par.eval_device(vals.data_ptr(), stride, n, out.data_ptr(), stride, stream=s)
    off.eval_device(vals.data_ptr(), stride, n, out2.data_ptr(), stride, stream=s)
    torch.cuda.synchronize()
    assert torch.equal(out[:, :n], out2[:, :n])
    sel = np.unique(np.concatenate([np.arange(0, n, n // 37), [n - 2, n - 1]]))
    host = ff.synthetic_points(8, n, p, seed=5)[:, sel]
    assert np.array_equal(out[:, :n].cpu().numpy().view(np.uint64)[:, sel], orc.eval(np.ascontiguousarray(host)))
    if not generic:
        pts = ff.synthetic_points(8, 3000, p, seed=1)
        want = orc.eval(np.ascontiguousarray(pts[:, :40]))
        for k, th in [(1, 128), (2, 256), (4, 512), (8, 128), (8, 256), (8, 1024)]:
            par.set_launch(k, th)
            assert np.array_equal(par.evaluate_pre(pts)[:, :40], want), (k, th)
    par.close()
    off.close()


@pytest.mark.gpu
@pytest.mark.parametrize("generic", [False, True])
def test_long_sums_on_device(ctx, oracle, generic):
    """long sums are cut into partial sums that the rows' tapes evaluate in parallel chunks (two levels of shared rows,
    three kernels per launch): a single large function at the Reconstructor's bunch sizes and at a big batch, against
    the oracle and against the same collection compiled in one piece"""
    rng = random.Random(23)
    vars_ = ["x", "y", "z", "w"]
    pool = [f"{rng.randint(1, 10**12)}*x^{rng.randint(1, 30)}*y^{rng.randint(0, 9)}*z^{rng.randint(0, 5)}*w^{rng.randint(0, 2)}" for _ in range(2500)]
    num = "".join(rng.choice("+-") + t for t in pool)[1:]
    den = "+".join(pool[:600])
    funs = [f"({num})/({den})", f"(-{'-'.join(pool[:350])})/(x-y)", "x^30*y+z", f"({den})^2*3"]
    p = 4611686018427400249 if generic else ff.primes()[13]
    par = ff.GpuParser(ctx, funs, vars_)
    off = ff.GpuParser(ctx, funs, vars_, split=False)
    par.set_prime(p)
    off.set_prime(p)
    assert par.n_split_rows >= 10 and off.n_split_rows == 0 and par.cse_fallbacks == 0 and len(par.table_levels()) == 2
    orc = oracle.parse(funs, vars_, p)
    for n in (1, 2, 7, 128, 129, 3001):
        pts = ff.synthetic_points(4, n, p, seed=n)
        if n > 2:
            pts[:, 1] = 0
            pts[:2, 2] = 5                    # x - y = 0
        got = par.evaluate_pre(pts)
        sel = np.unique(np.concatenate([np.arange(min(n, 8)), np.arange(0, n, max(1, n // 7)), [n - 1]]))
        assert np.array_equal(got[:, sel], orc.eval(np.ascontiguousarray(pts[:, sel]))), n
        assert np.array_equal(got, off.evaluate_pre(pts)), n
    # the example collection of the reference (one function of 1900 instructions among 19)
    g = load_golden("s_y_4_v")
    sp = ff.GpuParser(ctx, g["functions"], g["vars"])
    sp.set_prime(p)
    assert sp.n_split_rows >= 4
    pts = ff.synthetic_points(4, 33, p, seed=2)
    assert np.array_equal(sp.evaluate_pre(pts), oracle.parse(g["functions"], g["vars"], p).eval(pts))
    sp.close()
    par.close()
    off.close()


@pytest.mark.gpu
def test_randomised_collections_through_the_kernels():
    """twenty seconds of scripts/fuzz_gpu.py: random collections of every kind compiled with random passes, evaluated on
    the GPU right after ffgpu_set_prime at batch sizes 1..3000 in two primes, every value against the oracle (the long
    campaigns are in DESIGN.md; this keeps the harness and the compile -> upload -> launch sequence under test)"""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fuzz_gpu.py"), "--seconds", "20", "--seed", "2026"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " 0 failure(s)" in r.stdout
