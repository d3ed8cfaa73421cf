"""The C++ headers of the boundary at run time, and the three documented divergences from the reference
(DESIGN.md section 2), each pinned to exactly its input class."""
import os
import subprocess

import numpy as np
import pytest

import firefly_b200 as ff
from conftest import ROOT

pytestmark = pytest.mark.gpu
P0 = 9223372036854775783


def _splitmix(s, p):
    s = (s + 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
    z = s
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2 ** 64 - 1)
    z ^= z >> 31
    return s, 1 + z % (p - 1)


def test_probe_queue_header_runs(oracle, tmp_path):
    """include/firefly_gpu/GpuProbeQueue.hpp, compiled as a user would and run: single probes and a line of probes go
    in, (index, values) come out -- every value against the oracle, indices complete and in Reconstructor's order"""
    exe = tmp_path / "probe_queue_main"
    lib = os.path.join(ROOT, "firefly_b200", "lib")
    subprocess.run(["g++", "-std=c++14", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "probe_queue_main.cpp"),
                    "-o", str(exe), "-L", lib, "-lffgpu", f"-Wl,-rpath,{lib}", "-pthread"], check=True)
    p = ff.primes()[3]
    n_single, n_line = 301, 1000
    out = subprocess.run([str(exe), str(p), str(n_single), str(n_line)], capture_output=True, text=True, check=True).stdout
    rows = [[int(x) for x in ln.split()] for ln in out.splitlines()]
    assert [r[0] for r in rows] == list(range(n_single + n_line))
    s, pts = 42, []
    for _ in range(n_single):
        pt = []
        for _ in range(3):
            s, v = _splitmix(s, p)
            pt.append(v)
        pts.append(pt)
    a, sh = [1], []
    for _ in range(2):
        s, v = _splitmix(s, p)
        a.append(v)
    for _ in range(3):
        s, v = _splitmix(s, p)
        sh.append(v)
    for _ in range(n_line):
        s, t = _splitmix(s, p)
        pts.append([(a[v] * t + sh[v]) % p for v in range(3)])
    funs = ["(x+y)^3/(z-x*y)", "x^7*y-3*z", "x/(y-y)", "12345678901234567890123*z^2+x"]
    want = oracle.parse(funs, ["x", "y", "z"], p).eval(np.array(pts, dtype=np.uint64).T.copy())
    got = np.array([r[1:] for r in rows], dtype=np.uint64).T
    assert np.array_equal(got, want)


# ---- the documented divergences ------------------------------------------------------------------
def test_divergence_lane_exponent_in_a_bunch(ctx, oracle, ref):
    """(i) x^(-y) in a bunch: the reference's FFIntVec::to_neg_int reads lane 0 only (FFIntVec.cpp:48-51), so every lane
    of a bunch gets lane 0's exponent; the scalar path gives each probe its own.  The GPU is the scalar path for every
    batch size -- and agrees with the reference's bunch path exactly where that one agrees with itself (all lanes share
    the exponent), and for every function that has no lane-dependent negative power"""
    vars_ = ["x", "y"]
    p = ff.primes()[0]
    ref.set_prime(p)
    lane, plain = ["x^(-y)", "(x+1)^(-y)*y"], ["x^(-3)+y", "x^y", "(x*y)^(-2)"]
    pts = np.array([[5, 7, 11, 13], [2, 3, p - 2, 4]], dtype=np.uint64)          # four different exponents
    same = pts.copy()
    same[1, :] = 3                                                              # all lanes share the exponent
    for funs, differs in ((lane, True), (plain, False)):
        par = ff.GpuParser(ctx, funs, vars_)
        par.set_prime(p)
        rp = ref.parse(funs, vars_)
        scalar = rp.eval(pts)                                                   # evaluate_pre<FFInt>, point by point
        assert np.array_equal(par.evaluate_pre(pts), scalar) and np.array_equal(oracle.parse(funs, vars_, p).eval(pts), scalar)
        bunched = rp.eval(pts, bunch=4)                                         # evaluate_pre<FFIntVec<4>>
        assert (not np.array_equal(bunched, scalar)) == differs, funs
        assert np.array_equal(bunched[:, 0], scalar[:, 0])                      # lane 0 is always right
        assert np.array_equal(par.evaluate_pre(same), rp.eval(same, bunch=4))
        par.close()
        rp.close()


def test_divergence_constant_quotient_after_prime_change(ctx, oracle, ref):
    """(ii) x/(a/b): after a prime change the reference's token keeps only `a` (ShuntingYardParser.cpp:641), a freshly
    parsed reference and the GPU recompute the constant; functions without a quotient of constants are identical after
    prime_changed() as well"""
    vars_ = ["x"]
    p1, p2 = ff.primes()[0], ff.primes()[1]
    pts = np.array([[3, 10, 12345]], dtype=np.uint64)
    for funs, differs in ((["x/(3/7)", "x*(5/9)+1"], True), (["x/3+x*7", "(x+5)/(x-9)"], False)):
        ref.set_prime(p1)
        rp = ref.parse(funs, vars_)
        par = ff.GpuParser(ctx, funs, vars_)
        par.set_prime(p1)
        assert np.array_equal(par.evaluate_pre(pts), rp.eval(pts))
        ref.set_prime(p2)
        rp.prime_changed()
        stale = rp.eval(pts)
        fresh_parser = ref.parse(funs, vars_)
        fresh = fresh_parser.eval(pts)
        par.set_prime(p2)
        got = par.evaluate_pre(pts)
        assert np.array_equal(got, fresh) and np.array_equal(got, oracle.parse(funs, vars_, p2).eval(pts))
        assert (not np.array_equal(stale, fresh)) == differs, funs
        par.close()
        rp.close()
        fresh_parser.close()
    ref.set_prime(p1)


def test_divergence_undefined_inputs_are_rejected(ctx, oracle):
    """(iii) an operator without operands pops an empty std::stack in the reference (undefined behaviour); here it is an
    error -- and only that: the neighbouring well-formed spellings compile and match the oracle"""
    vars_ = ["x", "y"]
    for bad in ("x*-y", "x++*y", "*x", "x^^2"):
        with pytest.raises(ff.FFGpuError):
            ff.GpuParser(ctx, [bad], vars_)
    good = ["x*(-y)", "x+-y", "-x", "x^2", "+x-(-y)"]
    p = ff.primes()[2]
    par = ff.GpuParser(ctx, good, vars_)
    par.set_prime(p)
    pts = ff.synthetic_points(2, 9, p, seed=8)
    assert np.array_equal(par.evaluate_pre(pts), oracle.parse(good, vars_, p).eval(pts))
    par.close()
