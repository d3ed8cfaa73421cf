"""Randomised campaign over add_otf / trim on compiled collections -- TEST INFRASTRUCTURE, runs without a GPU.

Every round compiles the head of a random collection, adds the rest with ffgpu_add_function (with and without
check_is_equal / no_duplicates), trims a random subset, adds again, rebuilds the tape with set_prime and compares every
value of the interpreted tape (tests/tape_interp.py) with the C oracle evaluating the list of functions that should
have survived.

    python scripts/fuzz_add_trim.py <seed> <seconds>
"""
import sys, random, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'scripts')):
    sys.path.insert(0, d)
import numpy as np
import firefly_b200 as ff
from oracle_bind import Oracle
from tape_interp import run
import fuzz_compiler as F
orc = Oracle()
t0=time.time(); rounds=0; bad=0
seed0=int(sys.argv[1])
while time.time()-t0 < float(sys.argv[2]):
    rng = random.Random(seed0*100000+rounds); rounds+=1
    nv = rng.randint(1,4); vars_=[f"x{i}" for i in range(1,nv+1)]
    kind = rng.choice(["table","free"])
    funs = F.table(rng, vars_, rng.randint(4,20)) if kind=="table" else [f for f in F.free_grammar(rng, vars_, rng.randint(4,20), 3)]
    ok=[]
    for f in funs:
        try: ff.GpuParser(None,[f],vars_).close(); ok.append(f)
        except ff.FFGpuError: pass
    funs=ok
    if len(funs)<3: continue
    chk = rng.random()<0.4 and not any(__import__('re').search(r"\^\((?!-?\d+\))", f) for f in funs)
    k = rng.randint(0,len(funs)-1)
    opts=dict(cse=rng.random()<0.8, fuse=rng.random()<0.8, cross=rng.random()<0.8, temps=rng.choice([0,0,2,5]))
    par = ff.GpuParser(None, funs[:k], vars_, chk, **opts)
    p = rng.choice(ff.primes())
    if rng.random()<0.5: par.set_prime(p)
    cur = list(funs[:k])       # expected OUTPUT list
    for f in funs[k:]:
        nd = chk and rng.random()<0.3
        before = par.n_out
        par.add_otf(f, nd)
        if par.n_out>before: cur.append(f)
    if rng.random()<0.7 and cur:
        keep = sorted(rng.sample(range(len(cur)), rng.randint(1,len(cur))))
        m = par.trim(keep)
        assert sorted(m.keys())==keep and [m[q] for q in keep]==list(range(len(keep))), (m, keep)
        cur=[cur[q] for q in keep]
        # more additions after the trim
        for f in rng.sample(funs, min(3,len(funs))):
            before = par.n_out
            par.add_otf(f, False)
            if par.n_out>before: cur.append(f)
    par.set_prime(p)
    assert par.n_out==len(cur), (par.n_out, len(cur))
    pts = ff.synthetic_points(nv, 4, p, seed=rounds); pts[:,0]=0; pts[:,1]=1
    want = orc.parse(cur, vars_, p).eval(pts)
    tape = par.tape()
    for j in range(4):
        got = run(tape, [int(x) for x in pts[:,j]], p, par.n_out)
        if got != [int(x) for x in want[:,j]]:
            bad+=1; print("FAIL", seed0, rounds-1, kind, chk, opts); break
    par.close()
print(rounds, "rounds", bad, "failures")
sys.exit(1 if bad else 0)
