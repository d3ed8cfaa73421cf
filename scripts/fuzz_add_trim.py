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
    head = 0 if chk else k                                  # without the flag the head goes through the constructor
    par = ff.GpuParser(None, funs[:head], vars_, chk, **opts)
    p = rng.choice(ff.primes())
    if rng.random()<0.5: par.set_prime(p)
    # Model of what the parser holds: the texts of its distinct functions and, per output, which of them it returns.
    # With check_is_equal the parser decides by VALUE at a random point which functions are the same (as the
    # reference does: x2/x2 and (..)^0 are "equal" although they differ at x2 = 0), so the model takes that decision
    # from the returned position and checks everything else: positions handed out, outputs kept by trim, the
    # renumbering, the values.
    texts, outs = list(funs[:head]), list(range(head))

    def model_add(f, pos, n_out_before):
        if pos == len(texts):
            texts.append(f)
        assert pos < len(texts), (pos, len(texts))
        if par.n_out > n_out_before:
            outs.append(pos)

    for f in funs[head:]:
        nd = chk and rng.random() < 0.3
        before = par.n_out
        pos = par.add_otf(f, nd)
        if not chk:
            assert pos == len(texts)
        model_add(f, pos, before)
        if rng.random() < 0.1:
            par.set_prime(p)                          # a rebuild in between
    if rng.random() < 0.7 and outs:
        keep = sorted(rng.sample(range(len(outs)), rng.randint(1, len(outs))))
        m = par.trim(keep)
        assert sorted(m.keys()) == keep and [m[q] for q in keep] == list(range(len(keep))), (m, keep)
        nt, no, copied = [], [], {}
        for i in keep:
            fn = outs[i]
            if fn not in copied:
                copied[fn] = len(nt)
                nt.append(texts[fn])
            no.append(copied[fn])
        texts, outs = nt, no
        assert par.n_functions == len(texts), (par.n_functions, len(texts))
        for f in rng.sample(funs, min(3, len(funs))):
            before = par.n_out
            model_add(f, par.add_otf(f, False), before)
    par.set_prime(p)
    assert par.n_out == len(outs) and par.n_functions == len(texts), (par.n_out, len(outs), par.n_functions, len(texts))
    cur = [texts[i] for i in outs]
    pts = ff.synthetic_points(nv, 4, p, seed=rounds); pts[:,0]=0; pts[:,1]=1
    want = orc.parse(cur, vars_, p).eval(pts)
    tape = par.tape()
    levels = par.table_levels()
    for j in range(4):
        pt = [int(x) for x in pts[:,j]]
        full = pt + [0] * par.n_table_rows          # shared rows: level by level, each at its offset (include/ffgpu.h)
        for pre, rows, off in levels:
            full[len(pt) + off:len(pt) + off + rows] = run(pre, full, p, rows)
        got = run(tape, full, p, par.n_out)
        if got != [int(x) for x in want[:,j]]:
            bad+=1; print("FAIL", seed0, rounds-1, kind, chk, opts); break
    par.close()
print(rounds, "rounds", bad, "failures")
sys.exit(1 if bad else 0)
