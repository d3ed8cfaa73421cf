"""What the FIRST ffgpu_eval of a process and of a caller slot costs (development measurement).

Short reconstructions (s_y_safe: 136 probes) report a black-box time per probe that is mostly this."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(sys.path[0], "tests"))
import numpy as np
t0 = time.perf_counter()
import firefly_b200 as ff
from conftest import load_golden
def ms(t): return f"{(time.perf_counter() - t) * 1e3:9.3f} ms"
print("import", ms(t0))
t = time.perf_counter(); ctx = ff.Context(); print("ffgpu_create          ", ms(t))
for name in ("s_y_safe", "s_y_4_v"):
    g = load_golden(name)
    t = time.perf_counter(); par = ff.GpuParser(ctx, g["functions"], g["vars"]); print(name, "compile      ", ms(t))
    t = time.perf_counter(); par.set_prime(ff.primes()[0]); print(name, "set_prime    ", ms(t))
    pts = ff.synthetic_points(len(g["vars"]), 1, ff.primes()[0], seed=3)
    for slot in (0, 0, 0, 1, 1, 2, 2):
        t = time.perf_counter(); par.evaluate_pre(pts, thread_id=slot); print(name, f"eval slot {slot}  ", ms(t))
    t = time.perf_counter(); par.set_prime(ff.primes()[1]); print(name, "set_prime 2  ", ms(t))
    t = time.perf_counter(); par.evaluate_pre(pts, thread_id=0); print(name, "eval slot 0  ", ms(t))
