"""Workload generators: the benchmark functions of BASELINE.json's configs, built from their
formulas (arXiv:2004.01463 Eqs. 18-20, 29, 36-40) rather than shipped as files, and the
synthetic IBP-like collection of config 4.

Reference inputs these reproduce token for token (checked by tests/test_compiler.py::test_workloads_equal_reference_files when the
reference tree is present): benchmarks/f1.m-f4.m, benchmarks/factors1.m-factors5.m
(reference: benchmarks/benchmarks.cpp:71-84,119-131).
"""
from __future__ import annotations

import random


def _vars(n):
    return [f"x{i}" for i in range(1, n + 1)]


def _psum(n, e, signs=None):
    terms = []
    for i in range(1, n + 1):
        s = "-" if signs and signs[i - 1] < 0 else "+"
        terms.append((s if i > 1 else "") + f"x{i}^{e}")
    return "".join(terms)


def benchmark(name: str):
    """-> (function text, variable names)"""
    lin5 = "(1+x1+x2+x3+x4+x5)"
    lin4 = "(1+x1+x2+x3+x4)"
    prod5 = "x1*x2*x3*x4*x5"
    q = "(x1*x2+x3*x4+x5*x6)"
    qsum = f"({q}+{q}^2+{q}^3+{q}^4+{q}^5)"
    if name == "f1":      # Eq. (18): 20 variables
        return f"({_psum(20, 20)})/({_psum(20, 20, [1] * 10 + [-1] * 10)})", _vars(20)
    if name == "f2":      # Eq. (19)
        return f"(x1^100+x2^200+x3^300)/({prod5}+({prod5})^4)", _vars(5)
    if name == "f3":      # Eq. (20)
        return f"({lin5}^17-1)/(x4-x2+({prod5})^10)", _vars(5)
    if name == "f4":      # Eq. (29)
        return f"({lin5}^20-1)/(x4-x2+({prod5})^10)", _vars(5)
    if name == "factors1":
        return f"({lin4}^20-1)/({lin4}^20-1+x1^20)", _vars(4)
    if name == "factors2":
        return f"(x4-4)*(x4-3)*({lin4}^20-1)/({lin4}^20-1+x1^20)", _vars(4)
    if name == "factors3":
        return f"({_psum(20, 20)})/({qsum})", _vars(20)
    if name == "factors4":
        return f"({_psum(20, 20)})/({qsum}*x20^35)", _vars(20)
    if name == "factors5":
        return f"({_psum(20, 20)})/({qsum}*(x2*x20^35-1))", _vars(20)
    raise KeyError(name)


BENCHMARKS = ["f1", "f2", "f3", "f4", "factors1", "factors2", "factors3", "factors4", "factors5"]


def ibp_like_collection(n_functions: int = 10000, n_vars: int = 8, seed: int = 20200401,
                        min_terms: int = 3, max_terms: int = 24, den_run: int = 0):
    """Config 4: random sparse rational functions shaped like IBP reduction coefficients --
    Horner-nested polynomials in a few of the variables with integer and rational coefficients,
    exponents up to 30, at most one division each (numerator / denominator).
    -> (list of function strings, variable names).  Deterministic in (n_functions, n_vars, seed, den_run).
    den_run > 0: the coefficients come in runs of 1..den_run consecutive functions over the SAME denominator, the
    way the coefficients of one integral share their denominator in a reduction table (den_run = 0: every
    function has its own random denominator -- the collection all earlier measurements used)."""
    rng = random.Random(seed)
    names = [f"x{i}" for i in range(1, n_vars + 1)]

    def coeff():
        r = rng.random()
        if r < 0.55:
            return str(rng.randint(1, 10 ** rng.randint(1, 12)))
        if r < 0.85:
            return f"{rng.randint(1, 10 ** 9)}/{rng.randint(2, 10 ** 6)}"
        return str(rng.randint(10 ** 18, 10 ** 24))

    def horner(vars_, depth):
        # c0 + v*(c1 + v*(...)) with sub-polynomials in the remaining variables as coefficients
        v = vars_[0]
        rest = vars_[1:]
        n = rng.randint(1, 4)
        expr = None
        for _ in range(n):
            c = coeff() if (not rest or depth <= 0 or rng.random() < 0.5) else "(" + horner(rest, depth - 1) + ")"
            sign = "-" if rng.random() < 0.4 else "+"
            if expr is None:
                expr = c
            else:
                e = rng.choice([1, 1, 1, 2, 2, 3, 5, 8, 13, 30])
                pw = v if e == 1 else f"{v}^{e}"
                expr = f"{c}{sign}{pw}*({expr})"
        return expr

    def poly(target_terms):
        k = rng.randint(1, min(4, n_vars))
        vs = rng.sample(names, k)
        parts = []
        for _ in range(max(1, target_terms // 6)):
            rng.shuffle(vs)
            parts.append(horner(list(vs), 2))
        return "+".join(f"({p})" for p in parts)

    funs = []
    left, den = 0, None
    for _ in range(n_functions):
        t = rng.randint(min_terms, max_terms)
        num = poly(t)
        if den_run > 0:
            if left == 0:
                left = rng.randint(1, den_run)
                den = poly(max(min_terms, t // 2)) if rng.random() < 0.9 else None
            left -= 1
            funs.append(f"({num})/({den})" if den is not None else num)
        elif rng.random() < 0.8:
            den = poly(max(min_terms, t // 2))
            funs.append(f"({num})/({den})")
        else:
            funs.append(num)
    return funs, names
