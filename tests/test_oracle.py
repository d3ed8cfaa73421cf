"""The oracle (oracle/ff_oracle.c) pinned against the golden vectors and the unmodified reference."""
import random

import numpy as np
import pytest

from conftest import golden_names, load_golden

P0 = 9223372036854775783


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_golden(oracle, name):
    g = load_golden(name)
    for case in g["cases"]:
        par = oracle.parse(g["functions"], g["vars"], case["prime"])
        assert [par.rpn(i) for i in range(par.n_functions)] == g["rpn"]
        got = par.eval(case["points"])
        assert np.array_equal(got, case["values"]), name
        par.close()


def test_oracle_prime_change_matches_fresh(oracle):
    g = load_golden("s_y_4_v")
    par = oracle.parse(g["functions"], g["vars"], g["cases"][0]["prime"])
    for case in g["cases"][1:] + g["cases"][:1]:
        par.set_prime(case["prime"])
        assert np.array_equal(par.eval(case["points"]), case["values"])


def test_oracle_text_mode_splits_like_a_file(oracle):
    text = " x + 1 ;\n\n y*\t2;;  ;x^2\n"
    par = oracle.parse_text(text, ["x", "y"], P0)
    assert par.n_functions == 3
    assert par.rpn(0) == ["x", "1", "+"] and par.rpn(2) == ["x", "2", "^"]


def test_oracle_errors(oracle):
    for bad in ["(x", "x)", "q+1", "x*", ""]:
        with pytest.raises(ValueError):
            oracle.parse([bad], ["x"], P0).eval(np.array([[1]], dtype=np.uint64))


def test_oracle_duplicates(oracle):
    funs = ["x+y", "y+x", "x*y", "(x+y)", "x*y*1"]
    par = oracle.parse(funs, ["x", "y"], P0, check_is_equal=True)
    assert par.n_functions == 2 and par.n_out == 5
    v = par.eval(np.array([[3], [4]], dtype=np.uint64))[:, 0].tolist()
    assert v == [7, 7, 12, 7, 12]


def test_oracle_duplicates_vs_reference(oracle, ref):
    """check_is_equal goes through the string evaluator ShuntingYardParser.cpp:379-479, which knows the
    negative-exponent operators '~' and ';' (the public template evaluate<T>(values) does not): same set of
    distinct functions, same evaluation positions, same values as the reference"""
    import firefly_b200 as ff
    rng = random.Random(23)
    vars_ = ["x1", "x2", "x3"]
    base = ["x3^(-7)", "(-x1)^(-3)", "1/x3^7", "(x3)^(x2)*7/(x2-2733)^(-x3)", "x2^(-x1)", "-x2^(-2)+x1", "x1*x2-x3",
            "x3*(-x2)^3", "(x1+x2)^(-2)*x3", "x1^(-x2^2)", "12345678901234567890123*x1^(-1)"]
    funs = [rng.choice(base) for _ in range(40)] + ["x3^(-7)+0", "(x1*x2-x3)", "x2*x1-x3"]
    for k in (0, 7):
        p = ref.lib.ffref_prime(k)
        ref.set_prime(p)
        rp = ref.parse(funs, vars_, check_is_equal=True)
        op = oracle.parse(funs, vars_, p, check_is_equal=True)
        assert rp.n_out == op.n_out == len(funs) and rp.n_functions == op.n_functions < len(funs)
        assert [op.rpn(i) for i in range(op.n_functions)] == [rp.rpn(i) for i in range(rp.n_functions)]
        pts = ff.synthetic_points(3, 12, p, seed=k)
        pts[:, 0] = 0
        pts[:, 1] = np.array([1, p - 1, 2], dtype=np.uint64)
        assert np.array_equal(op.eval(pts), rp.eval(pts))
    ref.set_prime(ref.lib.ffref_prime(0))


def test_primes_table(oracle):
    L = oracle.lib
    ps = [L.ffo_prime(i) for i in range(300)]
    assert ps[0] == P0 and ps[1] == 9223372036854775643
    assert all(a > b for a, b in zip(ps, ps[1:]))
    assert all(25 <= (1 << 63) - p <= 13797 for p in ps)
    assert (1 << 63) - ps[-1] == 13797


# ---- against the unmodified reference ------------------------------------------------------------
def test_oracle_arithmetic_vs_reference(oracle, ref):
    rng = random.Random(7)
    L, R = oracle.lib, ref.lib
    for p in (P0, R.ffref_prime(299)):
        ref.set_prime(p)
        edge = [0, 1, 2, p - 1, p - 2, p >> 1, (p >> 1) + 1, (p >> 1) - 1, 0xFFFFFFFF, 0xFFFFFFFE, 0x80000000, 0x7FFFFFFF]
        vals = edge + [rng.randrange(p) for _ in range(300)]
        for _ in range(3000):
            a, b = rng.choice(vals), rng.choice(vals)
            assert L.ffo_add(a, b, p) == R.ffref_add(a, b)
            assert L.ffo_sub(a, b, p) == R.ffref_sub(a, b)
            assert L.ffo_mul(a, b, p) == R.ffref_mul(a, b)
            assert L.ffo_div(a, b, p) == R.ffref_div(a, b)
            assert L.ffo_neg(a, p) == R.ffref_neg(a)
            assert L.ffo_to_neg_int(a) == R.ffref_to_neg_int(a)
        for _ in range(300):
            a, e = rng.choice(vals), rng.choice(vals)
            assert L.ffo_pow(a, e, p) == R.ffref_pow(a, e)
            ei = rng.choice([0, 1, 2, 3, -1, -2, -3, 17, -17, 2**31 - 1, -(2**31 - 1), rng.randrange(-10**6, 10**6)])
            assert L.ffo_pow_int(a, ei, p) == R.ffref_pow_int(a, ei)


def test_bunch_policy_vs_reference(oracle, ref):
    L, R = oracle.lib, ref.lib
    for q in list(range(0, 300)) + [1000, 4096, 65537]:
        for thr in (1, 2, 3, 4, 7, 8, 48):
            for mb in (1, 2, 4, 8, 16, 32, 64, 128):
                assert L.ffo_compute_bunch_size(q, thr, mb) == R.ffref_compute_bunch_size(q, thr, mb)
                if q:
                    assert L.ffo_compute_job_number(q, thr, thr, mb) == R.ffref_compute_job_number(q, thr, thr, mb)


def test_primes_vs_reference(oracle, ref):
    assert [oracle.lib.ffo_prime(i) for i in range(300)] == [ref.lib.ffref_prime(i) for i in range(300)]


def test_oracle_random_expressions_vs_reference(oracle, ref):
    """Random expressions through both parsers: same RPN, same values, in two primes, with a prime change."""
    rng = random.Random(11)
    vars_ = ["a", "b", "c1"]

    def atom():
        r = rng.random()
        if r < 0.45:
            return rng.choice(vars_)
        if r < 0.8:
            return str(rng.randint(0, 10 ** rng.randint(1, 17)))
        return str(rng.randint(10 ** 18, 10 ** 25))

    def expr(d):
        if d == 0 or rng.random() < 0.25:
            return atom()
        r = rng.random()
        if r < 0.55:
            return f"{expr(d-1)}{rng.choice('+-*/')}{expr(d-1)}"
        if r < 0.7:
            return f"({expr(d-1)})"
        if r < 0.8:
            return f"({expr(d-1)})^{rng.randint(0, 40)}"
        if r < 0.87:
            return f"{atom()}^(-{rng.randint(1, 9)})"
        if r < 0.93:
            return f"(-{atom()})^{rng.randint(1, 5)}"
        if r < 0.97:
            return f"(-({expr(d-1)}))"
        return f"(-{rng.choice(vars_)}^{rng.randint(2, 7)}+{expr(d-1)})"

    funs = [expr(4) for _ in range(200)]
    p0, p1 = ref.lib.ffref_prime(0), ref.lib.ffref_prime(5)
    ref.set_prime(p0)
    rp = ref.parse(funs, vars_)
    op = oracle.parse(funs, vars_, p0)
    assert [op.rpn(i) for i in range(len(funs))] == [rp.rpn(i) for i in range(len(funs))]
    import firefly_b200 as ff
    pts = ff.synthetic_points(3, 16, p0, seed=5)
    pts[:, 0] = 0
    assert np.array_equal(op.eval(pts), rp.eval(pts))
    # the bunched reference path agrees with the scalar one on these inputs
    assert np.array_equal(rp.eval(pts, bunch=8, threads=2), rp.eval(pts))
    ref.set_prime(p1)
    rp.prime_changed()
    op.set_prime(p1)
    pts = ff.synthetic_points(3, 16, p1, seed=6)
    assert np.array_equal(op.eval(pts), rp.eval(pts))
    ref.set_prime(p0)
