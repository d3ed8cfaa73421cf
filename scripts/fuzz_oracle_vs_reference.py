"""Randomised campaign: the C oracle (oracle/ff_oracle.c) against the UNMODIFIED reference (oracle/_ref) -- TEST
INFRASTRUCTURE, runs without a GPU, needs `make -C oracle ref`.

Random collections from the grammar of scripts/fuzz_compiler.py (inputs the reference cannot evaluate -- an operator
without operands pops an empty std::stack there -- are dropped first, using the product's compiler as the filter);
RPN strings and values of evaluate_pre<FFInt> and evaluate_pre<FFIntVec<4>> are compared on edge and random points
for two primes, the second one reached through precompute_tokens() on the same parser -- unless a function holds a
quotient of constants, as in `x/(a/b)` or `x/((a/b)/c)`: the reference's recipe for that fold keeps only `a` after a
prime change (ShuntingYardParser.cpp:641,
DESIGN.md section 2 (ii)), so those collections are parsed afresh for the second prime.  Variable exponents are left
out of the bunched comparison (FFIntVec::to_neg_int reads lane 0 only).

    python scripts/fuzz_oracle_vs_reference.py <seed> <seconds>
"""
import os
import random
import re
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "scripts")):
    sys.path.insert(0, d)

import numpy as np  # noqa: E402

import firefly_b200 as ff  # noqa: E402
import fuzz_compiler as F  # noqa: E402
from oracle_bind import Oracle  # noqa: E402
from refs import Ref  # noqa: E402


def main():
    seed0, seconds = int(sys.argv[1]), float(sys.argv[2])
    orc, ref = Oracle(), Ref()
    primes = ff.primes()
    t0 = time.time()
    rounds = values = bad = 0
    while time.time() - t0 < seconds:
        rng = random.Random(seed0 * 1000003 + rounds)
        rounds += 1
        nv = rng.randint(1, 5)
        vars_ = [f"x{i}" for i in range(1, nv + 1)]
        funs = F.free_grammar(rng, vars_, rng.randint(1, 25), rng.randint(1, 5)) if rng.random() < 0.7 \
            else F.table(rng, vars_, rng.randint(2, 25))
        kept = []
        for f in funs:
            try:
                ff.GpuParser(None, [f], vars_).close()
                kept.append(f)
            except ff.FFGpuError:
                pass
        funs = kept
        if not funs:
            continue
        lane_exponents = any(re.search(r"\^\((?!-?\d+\))", f) for f in funs)
        p1, p2 = rng.choice(primes), rng.choice(primes)
        ref.set_prime(p1)
        rp = ref.parse(funs, vars_)
        def is_num(t):
            return t.lstrip("+-")[:1].isdigit()

        rpns = [rp.rpn(i) for i in range(len(funs))]

        def has_constant_quotient(r):
            st = []                                    # True = constant sub-expression
            for t in r:
                if len(t) == 1 and t in "+-*/^!~;":
                    b, a = st.pop(), st.pop()
                    if t == "/" and a and b:
                        return True
                    st.append(a and b)
                else:
                    st.append(is_num(t))
            return False

        stale_fold = any(has_constant_quotient(r) for r in rpns)
        for k, p in enumerate((p1, p2)):
            if k:
                ref.set_prime(p)
                if stale_fold:
                    rp.close()
                    rp = ref.parse(funs, vars_)
                else:
                    rp.prime_changed()
            op = orc.parse(funs, vars_, p)
            if [op.rpn(i) for i in range(len(funs))] != [rp.rpn(i) for i in range(len(funs))]:
                bad += 1
                print("FAIL rpn", seed0, rounds - 1, flush=True)
                break
            pts = ff.synthetic_points(nv, 8, p, seed=rng.randrange(1 << 60))
            pts[:, 0] = 0
            pts[:, 1] = 1
            pts[:, 2] = p - 1
            pts[:, 3] = pts[0, 3]
            want = rp.eval(pts)
            got = op.eval(pts)
            if not np.array_equal(got, want) or (not lane_exponents and not np.array_equal(rp.eval(pts, bunch=4), want)):
                bad += 1
                j, i = np.argwhere(got != want)[0][::-1] if not np.array_equal(got, want) else (0, 0)
                print("FAIL value", seed0, rounds - 1, "p", p, "point", j, "function", funs[i], flush=True)
                break
            values += want.size
            op.close()
        rp.close()
    print(f"{rounds} collections, {values} values identical, {bad} failure(s)")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
