"""Any odd prime below 2^63 on the device (north_star: "Montgomery/Barrett preinverse"; reference: FFInt::set_new_prime
takes any modulus, source/FFInt.cpp:344-353, and its mod_mul / mod_pow / mod_inv are generic, :356-405).  Primes of the
form 2^63 - c, c < 2^15 -- all of FireFly's table -- keep the tuned arithmetic; every other prime runs on the Barrett
kernels.  Bit-exact against Python integers (field operations) and the oracle (whole collections)."""
import random

import numpy as np
import pytest

import firefly_b200 as ff
from conftest import load_golden
from firefly_b200 import workloads as W

pytestmark = pytest.mark.gpu

# just outside the table form (c = 32835 >= 2^15), 2^63 - 2^20 - .., 62 bits, the Mersenne prime 2^61 - 1, 40 bits,
# just above 2^32, three digits
PRIMES = [9223372036854742973, 9223372036853727157, 4611686018427400249, 2305843009213693951, 1099511627689, 4294967311, 997, 103]


def _inv(a, p):
    return pow(a, p - 2, p) if a % p else 0


def _pow_ff(b, e, p):
    if e == 2:
        return b * b % p
    if e < (p >> 1):
        return pow(b, e, p)
    return pow(_inv(b, p), p - e - 1, p)


def _to_neg_int(n):
    v = (-(n & 0xFFFFFFFF)) & 0xFFFFFFFF
    return v - (1 << 32) if v >= (1 << 31) else v


def _pow_int(b, e, p):
    if e == 2:
        return b * b % p
    return pow(b, e, p) if e >= 0 else pow(_inv(b, p), -e, p)


@pytest.mark.parametrize("p", PRIMES)
def test_field_ops_generic_modulus(ctx, p):
    rng = random.Random(p)
    edge = [0, 1, 2, p - 1, p - 2, p >> 1, (p >> 1) + 1, p, p + 1, 2 * p + 1, (1 << 64) - 1, (1 << 63), 12345678901234567890 % (1 << 64)]
    n = 20000
    a = np.array([rng.choice(edge) if i % 5 == 0 else rng.randrange(p) for i in range(n)], dtype=np.uint64)
    b = np.array([rng.choice(edge) if i % 7 == 0 else rng.randrange(p) for i in range(n)], dtype=np.uint64)
    ne = len(edge)
    a[: ne * ne] = np.repeat(np.array(edge, dtype=np.uint64), ne)
    b[: ne * ne] = np.tile(np.array(edge, dtype=np.uint64), ne)
    ai, bi = [int(x) % p for x in a], [int(x) % p for x in b]       # non-residues are reduced on the way in
    cases = [(0, lambda x, y: (x + y) % p), (1, lambda x, y: (x - y) % p), (2, lambda x, y: x * y % p),
             (3, lambda x, y: (-x) % p), (4, lambda x, y: _inv(x, p)), (5, lambda x, y: _pow_ff(x, y, p)),
             (6, lambda x, y: _inv(x, p)), (7, lambda x, y: _pow_int(x, _to_neg_int(y), p)), (8, lambda x, y: x % p),
             (9, lambda x, y: _inv(x, p) * _inv(y, p) % p),
             (10, lambda x, y: ((x + y) * y) % p), (11, lambda x, y: ((x - y) * y) % p),
             (12, lambda x, y: ((x * y) * (x * y) - y) % p), (13, lambda x, y: ((-x) * y + x) % p)]
    for op, fn in cases:
        m = 2000 if op in (5, 7) else n                     # lane powers: a divergent loop per value
        got = ctx.field_ops(op, a[:m], b[:m], p)
        want = np.array([fn(x, y) for x, y in zip(ai[:m], bi[:m])], dtype=np.uint64)
        assert np.array_equal(got, want), f"op {op} prime {p}"


@pytest.mark.parametrize("name", ["s_y_4_v", "edge_cases", "s_y_1_v", "s_y_safe", "f1", "f4", "factors4"])
@pytest.mark.parametrize("p", PRIMES[:5])
def test_collections_generic_modulus(ctx, oracle, name, p):
    g = load_golden(name)
    par = ff.GpuParser(ctx, g["functions"], g["vars"])
    par.set_prime(p)
    op = oracle.parse(g["functions"], g["vars"], p)
    for k, threads in [(0, 0), (1, 128), (2, 256), (4, 128), (4, 256), (8, 256)]:      # (8, 256) falls back to 4 probes per thread
        par.set_launch(k, threads)
        for n in (1, 2, 33, 1000, 4099):
            pts = ff.synthetic_points(par.n_vars, n, p, seed=n + k)
            if n >= 3:
                pts[:, 1] = 0
                pts[0, 2] = p - 1
            assert np.array_equal(par.evaluate_pre(pts), op.eval(pts)), (name, p, n, k, threads)
    par.close()


def test_prime_change_between_arithmetics(ctx, oracle):
    """prime_changed(): table prime -> generic prime -> table prime on one handle, big batch through the chunked path"""
    funs, vars_ = W.ibp_like_collection(60, 8, seed=2)
    par = ff.GpuParser(ctx, funs, vars_)
    n = 70001
    for p in (ff.primes()[0], PRIMES[0], ff.primes()[7], PRIMES[3], ff.primes()[0]):
        par.set_prime(p)
        pts = ff.synthetic_points(8, n, p, seed=3)
        pts[:, 5] = 0
        got = par.evaluate_pre(pts)
        idx = np.arange(0, n, 97)
        assert np.array_equal(got[:, idx], oracle.parse(funs, vars_, p).eval(np.ascontiguousarray(pts[:, idx]))), p


def test_eval_line_generic_modulus(ctx, oracle):
    text, vars_ = W.benchmark("f3")
    p = PRIMES[1]
    par = ff.GpuParser(ctx, [text], vars_)
    par.set_prime(p)
    rng = np.random.default_rng(4)
    a = rng.integers(1, p - 1, size=5, dtype=np.uint64)
    s = rng.integers(0, p - 1, size=5, dtype=np.uint64)
    t = rng.integers(0, (1 << 63), size=3001, dtype=np.uint64)          # some t are not residues: reduced on the device
    got = par.evaluate_line(a, s, t)
    pts = np.array([[(int(a[v]) * int(tj) + int(s[v])) % p for tj in t] for v in range(5)], dtype=np.uint64)
    assert np.array_equal(got, oracle.parse([text], vars_, p).eval(pts))


def test_rejected_moduli(ctx):
    par = ff.GpuParser(ctx, ["x+1"], ["x"])
    for bad in (0, 1, 2, 4, 10, 1 << 63, (1 << 63) + 1):
        with pytest.raises(ff.FFGpuError):
            par.set_prime(bad)


def test_literal_exponents_for_tiny_primes_on_device(ctx, oracle):
    """exponents at and above p/2 and above p (x^103 mod 103 = x^0 in the reference's arithmetic, FFInt.cpp:155-179): the
    same handle moves between tiny primes and FireFly's, the tape and its device copies are regenerated at the border"""
    funs = ["x^200+y", "x^60*y+x^60*3+x^61*y^2", "x^103", "5*x^104*y+7*x^105*y+x^104", "x^20+y^20+(x+y)^200", "(x^70-y)/(x^150+3)"]
    par = ff.GpuParser(ctx, funs, ["x", "y"])
    for p in (103, ff.primes()[0], 997, 103, 65537, 2097143, ff.primes()[3]):
        par.set_prime(p)
        for n in (1, 7, 300):
            pts = ff.synthetic_points(2, n, p, seed=n)
            pts[:, 0] = 0
            assert np.array_equal(par.evaluate_pre(pts), oracle.parse(funs, ["x", "y"], p).eval(pts)), (p, n)
    par.close()
