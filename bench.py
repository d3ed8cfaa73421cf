#!/usr/bin/env python
"""bench.py -- probe-evaluation throughput of the B200 path, one JSON line.

Workload of the headline line (BASELINE.json configs[1]): the paper's benchmark functions f1-f4 as black boxes.
One STEP = one pass of the hot path over one batch of synthetic probe points: each of f1, f2, f3, f4 evaluated at
`points` points mod p = 2^63 - 25 (one kernel launch per function).  Probes are independent, so N GPUs shard the
point index range (weak scaling: every rank owns `points` points per function); no collective on the data path.

  value         probes/s over all ranks with the points already resident in HBM (device API), CUDA events on the
                launching stream, barrier + synchronize on both sides, max over ranks
  e2e           the same metric through ffgpu_eval with HOST buffers (pinned), H2D and D2H inside the timed region
  e2e_line      ... through ffgpu_eval_line: only t crosses PCIe, the coordinates are formed on the GPU
  e2e_ctx       (N > 1, rank 0) ONE context over all N GPUs, one ffgpu_eval per function on a host batch that the
                library shards in-process by point index -- the path GpuBlackBox(..., devices) uses
  roofline      the interpreter kernel against the measured integer-multiply-pipe peak
                (profiles/r01_pipe_peak_v0.json; MEASURED_PEAKS.json has no integer peak) and HBM
  configs       the other BASELINE.json configurations, a few hundred ms each, every one checked against the oracle:
                cfg1 bunch sweep 1..4096 through the host API (+ 16 concurrent callers), cfg2 factors1-5 x
                primes()[0..7] sharded over the GPUs, cfg3 the 10 000-function IBP-like collection
                (cfg0 and cfg4 -- reconstructions through the reference's Reconstructor -- are tests:
                tests/test_gpu_reconstruct.py, log in gpurun_out/reconstruct_black_box_times.log)
  cpu_baseline  the reference's own evaluate_pre (oracle/_ref, FFIntVec<128>, all host threads) or the C port, on a
                bounded sample, rank 0 at N=1 only

`--impl reference` times the reference's CPU implementation on the same metric/config; it loads no GPU code.
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FUNCS = ["f1", "f2", "f3", "f4"]
FACTORS = ["factors1", "factors2", "factors3", "factors4", "factors5"]
P0 = 9223372036854775783
SEED = 0x4D595DF4D0F33173
METRIC = "mod_p_probe_evals_per_sec"
UNIT = "probes/s"
# multiply instructions per modular multiplication (4 product + 2 + 1 fold) and the
# charge for one inversion in modular multiplications
IMAD_PER_MM = 7
MM_PER_INV = 6
# pipe cycles: IMAD.WIDE issues at half the IMAD rate on B200 (profiles/r01_pipe_peak_v0.json), so the
# 6 wide + 1 narrow multiplies of one MM occupy 13 IMAD slots
SLOTS_PER_MM = 13
# device time of one step (f1-f4) per 2^24 points per function on one B200, for sizing the timed region
MS_PER_STEP_2P24 = 3.6


def _pure(name):
    """a module of firefly_b200/ that needs no library (workloads.py, points.py, shard.py), loaded WITHOUT the package:
    importing the package loads libffgpu.so, and the reference arm must not map any of the product's code"""
    spec = importlib.util.spec_from_file_location("ffpure_" + name, os.path.join(ROOT, "firefly_b200", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def measured_int_peak():
    """thread-instructions/s of 32-bit IMAD over the whole chip, measured by bench/pipe_peak.cu"""
    path = os.path.join(ROOT, "profiles", "r01_pipe_peak_v0.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return d["pipes"]["imad_lo"]["gthread_inst_per_s"] * 1e9, "profiles/r01_pipe_peak_v0.json imad_lo"
    except Exception:
        return 148 * 64 * 1.965e9, "nominal 148 SM x 64 IMAD/clk x 1.965 GHz"


def measured_modmul_peak():
    """G modular multiplications/s of the shipped mul_b63 / sqr_b63 chains alone (bench/modmul_forms.cu): what the
    arithmetic could do with no tape, no loads and no inversion around it"""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_modmul_forms_v1.json")) as f:
            d = json.load(f)
        return {"mul": d["mul_x1h0d0"]["gmm_per_s"], "sqr": d["sqr_x1h0d1"]["gmm_per_s"],
                "source": "profiles/r01_modmul_forms_v1.json (mul_x1h0d0, sqr_x1h0d1)"}
    except Exception:
        return None


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"], "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback"


def measured_link_peak(device_index=None):
    """pinned host <-> device copy rates (GB/s): build/h2d_peak (bench/h2d_peak.cu) run now on this rank's GPU -- under
    torchrun all ranks run it at the same time, so the figures are what the host gives N GPUs at once --, else the
    kept run under profiles/"""
    exe = os.path.join(ROOT, "build", "h2d_peak")
    if device_index is not None and os.path.exists(exe):
        try:
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(device_index))
            r = subprocess.run([exe, "--gpus", "1", "--mb", "128", "--reps", "5"], capture_output=True, text=True, env=env,
                               timeout=120, check=True)
            d = json.loads(r.stdout.strip().splitlines()[-1])
            return {"h2d_gbs": d["h2d_gbs"], "d2h_gbs": d["d2h_gbs"], "both_gbs_sum": d["both_gbs_sum"],
                    "source": "bench/h2d_peak.cu run inside this bench (128 MB pinned copies)"}
        except Exception:
            pass
    try:
        with open(os.path.join(ROOT, "profiles", "r02_h2d_peak.json")) as f:
            d = json.load(f)
        return {"h2d_gbs": d["h2d_gbs"], "d2h_gbs": d["d2h_gbs"], "both_gbs_sum": d["both_gbs_sum"],
                "source": "profiles/r02_h2d_peak.json (kept run)"}
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic(name):
    """DRAM bytes per probe (dram__bytes_read.sum + dram__bytes_write.sum over the launch's points) of the function's
    kernel from the committed `ncu --set full` capture, or None"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(name)
    except Exception:
        return None


def work_of(par):
    """(reference-token-list work, MM-eq of it, MM-eq the compiled tape executes)"""
    w = par.work()
    mm_eq = w["mm"] + MM_PER_INV * w["inv"]
    x = par.tape_work()
    w = dict(w, executed_mm=x["mm"], executed_inv=x["inv"], executed_ma=x["ma"], executed_var_reads=x["var_reads"],
             executed_mm_eq=x["mm"] + MM_PER_INV * x["inv"])
    return w, mm_eq


# ------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world, real_stdout):
    """The reference's own CPU evaluate_pre on this box's host cores (rank 0 only).  Loads oracle/ (the checker, here
    the thing timed: the one place besides cpu_baseline where bench.py may) and pure-Python helpers -- nothing of the
    product: libffgpu.so is not mapped by this process."""
    if rank != 0:
        return
    W = _pure("workloads")
    P = _pure("points")
    cores = os.cpu_count() or 1
    from refs import have_ref
    kind = "reference" if have_ref() else "port"
    n = args.cpu_points
    steps, warm = args.steps, args.warmup
    pars, pts = [], []
    if kind == "reference":
        from refs import Ref
        ref = Ref()
        ref.set_prime(P0)
        for name in FUNCS:
            text, vars_ = W.benchmark(name)
            pars.append(ref.parse([text], vars_))
            pts.append(P.synthetic_points(len(vars_), n, P0, seed=SEED))
        run = lambda i: pars[i].eval(pts[i], bunch=128, threads=cores)
        sample = f"{n} points per function, evaluate_pre<FFIntVec<128>>, {cores} threads"
    else:
        from oracle_bind import Oracle
        orc = Oracle()
        cores = 1
        for name in FUNCS:
            text, vars_ = W.benchmark(name)
            pars.append(orc.parse([text], vars_, P0))
            pts.append(P.synthetic_points(len(vars_), n, P0, seed=SEED))
        run = lambda i: pars[i].eval(pts[i])
        sample = f"{n} points per function, scalar C port, 1 thread"
    for _ in range(warm):
        for i in range(len(FUNCS)):
            run(i)
    t0 = time.perf_counter()
    for _ in range(steps):
        for i in range(len(FUNCS)):
            run(i)
    dt = time.perf_counter() - t0
    value = steps * len(FUNCS) * n / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "f1-f4 black-box probes (BASELINE.json configs[1])", "functions": FUNCS,
                   "prime": str(P0), "points_per_function_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "product_code_loaded": any("libffgpu" in ln for ln in open("/proc/self/maps")) if os.path.exists("/proc/self/maps") else None,
    }
    _emit(real_stdout, line)


# ------------------------------------------------------------------------------------------------------
def _claim_stdout():
    """stdout must carry exactly ONE JSON line: keep the real stdout for it and send everything else that any
    library writes to file descriptor 1 (NCCL prints its version banner there) to stderr"""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def _emit(real_stdout, line):
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


def pick_points(args):
    """points per function per step: a power of two between 2^24 and 2^28 that makes the timed region last about a
    second for the given number of steps (the step scales, not the step count)"""
    if args.points:
        return args.points
    want_ms = 1100.0 / max(args.steps, 1)
    n = 1 << 24
    while n < (1 << 28) and MS_PER_STEP_2P24 * (n >> 24) < want_ms:
        n <<= 1
    return n


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=0,
                    help="points per function per step per GPU (device-resident); 0 = sized for a timed region of ~1 s")
    ap.add_argument("--e2e-points", type=int, default=1 << 21, help="points per function per step per GPU (host buffers)")
    ap.add_argument("--cpu-points", type=int, default=1 << 20, help="CPU baseline sample: points per function")
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs block (cfg1 sweep, cfg2 factors, cfg3 IBP)")
    ap.add_argument("--ibp-functions", type=int, default=10000)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world, real_stdout)
        return

    import torch
    import torch.distributed as dist

    import firefly_b200 as ff
    from firefly_b200 import workloads as W
    from firefly_b200.shard import rank_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner / debug output goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(flag):
        t = torch.tensor([1 if flag else 0], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    from oracle_bind import Oracle
    orc = Oracle()
    ctx = ff.Context()
    n = pick_points(args)
    stride = n + (n & 1)
    stream = torch.cuda.current_stream().cuda_stream
    # this rank's shard of the global point index range (contiguous ranges by rank: firefly_b200/shard.py)
    first, _ = rank_range(n * world, rank, world)
    total_points = n * world
    pars, vals, outs, works = [], [], [], []
    for name in FUNCS:
        text, vars_ = W.benchmark(name)
        par = ff.GpuParser(ctx, [text], vars_)
        par.set_prime(P0)
        par.set_launch(args.k, args.threads)
        v = torch.empty((len(vars_), stride), dtype=torch.int64, device="cuda")
        o = torch.empty((1, stride), dtype=torch.int64, device="cuda")
        rc = ff.lib.ffgpu_fill_points(ctx.h, 0, v.data_ptr(), len(vars_), n, stride, first, total_points, SEED, P0, stream)
        assert rc == 0, ff.last_error()
        pars.append(par)
        vals.append(v)
        outs.append(o)
        works.append(work_of(par))
    torch.cuda.synchronize()

    def step():
        for par, v, o in zip(pars, vals, outs):
            par.eval_device(v.data_ptr(), stride, n, o.data_ptr(), stride, stream=stream)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: device-resident ------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(FUNCS) + 1)] for _ in range(args.steps)]
    launches0 = ff.launch_count()
    barrier()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for s in range(args.steps):
        ev[s][0].record()
        for i, (par, v, o) in enumerate(zip(pars, vals, outs)):
            par.eval_device(v.data_ptr(), stride, n, o.data_ptr(), stride, stream=stream)
            ev[s][i + 1].record()
    t_end.record()
    barrier()
    launches = ff.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = t_start.elapsed_time(t_end)
    per_fn_ms = [float(np.mean([ev[s][i].elapsed_time(ev[s][i + 1]) for s in range(args.steps)])) for i in range(len(FUNCS))]
    ms_total = max_over_ranks(ms_total)
    value = args.steps * len(FUNCS) * n * world / (ms_total * 1e-3)

    # ---- verification of what was just computed (rank-local sample against the oracle) ----------------------
    verified = True
    idx = np.unique(np.concatenate([np.arange(0, min(n, 2048)), np.arange(0, n, max(1, n // 2048)), [n - 1]]))
    for name, par, v, o in zip(FUNCS, pars, vals, outs):
        text, vars_ = W.benchmark(name)
        host_pts = v[:, torch.from_numpy(idx).cuda()].cpu().numpy().view(np.uint64)
        # the device generator is checked against its host twin on the first 4096 points of the shard
        m = min(n, 4096)
        twin = ff.synthetic_points(len(vars_), m, P0, seed=SEED, first_point=first, total_points=total_points)
        if not np.array_equal(v[:, :m].cpu().numpy().view(np.uint64), twin):
            verified = False
        want = orc.parse([text], vars_, P0).eval(np.ascontiguousarray(host_pts))
        got = o[:, torch.from_numpy(idx).cuda()].cpu().numpy().view(np.uint64)
        if not np.array_equal(got, want):
            verified = False
    verified = all_ok(verified)
    # the big device buffers are not needed any more
    del vals, outs
    torch.cuda.empty_cache()

    # ---- e2e: host buffers through the C ABI -------------------------------------------------------------------
    ne = args.e2e_points
    host_in, host_out, h2d, d2h = [], [], 0, 0
    import ctypes as C
    e_first, _ = rank_range(ne * world, rank, world)
    for name, par in zip(FUNCS, pars):
        nv = par.n_vars
        pin = ff.lib.ffgpu_host_alloc(nv * ne * 8)
        pout = ff.lib.ffgpu_host_alloc(ne * 8)
        assert pin and pout, ff.last_error()
        arr = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_uint64)), shape=(nv, ne))
        arr[:] = ff.synthetic_points(nv, ne, P0, seed=SEED + 1, first_point=e_first, total_points=ne * world)
        host_in.append((pin, arr))
        host_out.append((pout, np.ctypeslib.as_array(C.cast(pout, C.POINTER(C.c_uint64)), shape=(1, ne))))
        h2d += nv * ne * 8
        d2h += ne * 8

    def e2e_step():
        for par, (pin, _), (pout, _) in zip(pars, host_in, host_out):
            par.eval_ptr(pin, ne, pout)

    for _ in range(2):
        e2e_step()
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    e2e_l0 = ff.launch_count()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    e2e_launches = ff.launch_count() - e2e_l0
    e2e_value = e2e_steps * len(FUNCS) * ne * world / max_over_ranks(dt)
    # e2e results checked too
    for name, par, (pin, arr), (pout, oarr) in zip(FUNCS, pars, host_in, host_out):
        text, vars_ = W.benchmark(name)
        sel = np.arange(0, ne, max(1, ne // 1024))
        want = orc.parse([text], vars_, P0).eval(np.ascontiguousarray(arr[:, sel]))
        if not np.array_equal(oarr[:, sel], want):
            verified = False

    # ---- e2e, line form: only t crosses PCIe, x_v = a_v*t + s_v is formed on the GPU (what Reconstructor::queue_probes
    # computes per point, Reconstructor.hpp:2946-2960) --------------------------------------------------------------------
    rngl = np.random.default_rng(12345 + rank)
    line_in, line_out, line_as, lh2d, ld2h = [], [], [], 0, 0
    for name, par in zip(FUNCS, pars):
        nv = par.n_vars
        a = (rngl.integers(1, P0 - 1, size=nv, dtype=np.uint64))
        a[0] = 1
        sft = rngl.integers(0, P0 - 1, size=nv, dtype=np.uint64)
        pt = ff.lib.ffgpu_host_alloc(ne * 8)
        po = ff.lib.ffgpu_host_alloc(ne * 8)
        assert pt and po, ff.last_error()
        tarr = np.ctypeslib.as_array(C.cast(pt, C.POINTER(C.c_uint64)), shape=(ne,))
        tarr[:] = ff.synthetic_points(1, ne, P0, seed=SEED + 2, first_point=e_first, total_points=ne * world)[0]
        line_in.append((pt, tarr))
        line_out.append((po, np.ctypeslib.as_array(C.cast(po, C.POINTER(C.c_uint64)), shape=(1, ne))))
        line_as.append((a, sft))
        lh2d += ne * 8 + 2 * nv * 8
        ld2h += ne * 8

    def line_step():
        for par, (pt, _), (po, _), (a, sft) in zip(pars, line_in, line_out, line_as):
            par.eval_line_ptr(a, sft, pt, ne, po)

    for _ in range(2):
        line_step()
    barrier()
    line_l0 = ff.launch_count()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        line_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    line_launches = ff.launch_count() - line_l0
    line_value = e2e_steps * len(FUNCS) * ne * world / max_over_ranks(dt)
    for name, par, (pt, tarr), (po, oarr), (a, sft) in zip(FUNCS, pars, line_in, line_out, line_as):
        text, vars_ = W.benchmark(name)
        sel = np.arange(0, ne, max(1, ne // 512))
        pts = np.array([[(int(a[v]) * int(tj) + int(sft[v])) % P0 for tj in tarr[sel]] for v in range(par.n_vars)],
                       dtype=np.uint64)
        want = orc.parse([text], vars_, P0).eval(pts)
        if not np.array_equal(oarr[:, sel], want):
            verified = False
    verified = all_ok(verified)

    # ---- e2e_ctx: ONE context over all GPUs of the job, sharded in-process (rank 0; the other ranks wait) -------------
    e2e_ctx = None
    barrier()
    if world > 1:
        # the other ranks wait on the HOST (the job's key-value store): a collective's kernel must not sit on the GPUs
        # that rank 0's context is using meanwhile
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            try:
                e2e_ctx = run_e2e_ctx(ff, W, orc, world, ne, e2e_steps)
            finally:
                store.set("ffgpu_e2e_ctx_done", "1")
        else:
            store.wait(["ffgpu_e2e_ctx_done"])
    barrier()

    # ---- the other BASELINE configurations ----------------------------------------------------------------------------
    configs = None
    if not args.no_configs:
        configs = {}
        configs["cfg2_factors_x_primes"] = run_cfg_factors(ff, W, orc, ctx, torch, stream, rank, world, max_over_ranks, all_ok, barrier)
        configs["cfg3_ibp_like_collection"] = run_cfg_ibp(ff, W, orc, ctx, torch, stream, rank, world, max_over_ranks, all_ok, barrier,
                                                           args.ibp_functions)
        barrier()
        if rank == 0:
            configs["cfg1_bunch_sweep"] = run_cfg_sweep(ff, W, orc, ctx)
            configs["plugin_call_large_function"] = run_cfg_large_function(ff, orc, ctx)
        barrier()
        configs["cfg0_cfg4_reconstructions"] = (
            "tests/test_gpu_reconstruct.py: the unmodified reference Reconstructor over GpuBlackBox against its own CPU black "
            "box -- s_y_4_v (bunch 1 and 4), s_y_1_v, f1, factors3 and the full f4 run (bunch 128, 16 threads), identical "
            "results; the reference's average black-box time for both is logged to gpurun_out/reconstruct_black_box_times.log "
            "(kept run: profiles/r02_reconstruct_black_box_times.log)")
        for c in ("cfg2_factors_x_primes", "cfg3_ibp_like_collection"):
            verified = verified and configs[c]["verified_bit_exact_vs_oracle"]
        if rank == 0:
            verified = verified and configs["cfg1_bunch_sweep"]["verified_bit_exact_vs_oracle"]
            verified = verified and configs["plugin_call_large_function"]["verified_bit_exact_vs_oracle"]
        if e2e_ctx is not None:
            verified = verified and e2e_ctx["verified_bit_exact_vs_oracle"]

    # ---- what the host link gives this job: every rank copies at the same time ------------------------------------------
    barrier()
    physical = os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")
    dev_index = physical[local] if len(physical) > local and physical[local] else local
    link_own = measured_link_peak(dev_index)
    barrier()
    link = None
    if link_own:
        t = torch.tensor([link_own["h2d_gbs"], link_own["d2h_gbs"], link_own["both_gbs_sum"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        link = {"h2d_gbs": float(t[0]), "d2h_gbs": float(t[1]), "both_gbs_sum": float(t[2]), "gpus_copying_at_once": world,
                "source": link_own["source"]}

    # ---- CPU baseline (rank 0, N=1 only) ------------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline(args.cpu_points)
        except Exception as e:  # the baseline is a report, never a reason to lose the GPU numbers
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}

    if rank == 0:
        peak_int, peak_src = measured_int_peak()
        hbm_peak, hbm_src = measured_hbm_peak()
        # dominant kernel: the launch with the largest share of the step
        top = int(np.argmax(per_fn_ms))
        w, mm_eq = works[top]
        probes_per_s_top = n / (per_fn_ms[top] * 1e-3)
        imad = probes_per_s_top * mm_eq * IMAD_PER_MM
        bytes_per_probe = 8 * (pars[top].n_vars + 1)
        per_fn = {}
        for i, name in enumerate(FUNCS):
            wi, mmi = works[i]
            rate = n / (per_fn_ms[i] * 1e-3)
            per_fn[name] = {
                "ms": per_fn_ms[i], "probes_per_s": rate, "mm_eq_per_probe": mmi, "gmm_eq_per_s": rate * mmi / 1e9,
                "executed_mm_eq_per_probe": wi["executed_mm_eq"], "executed_gmm_eq_per_s": rate * wi["executed_mm_eq"] / 1e9,
                "imad_frac_of_measured": rate * mmi * IMAD_PER_MM / peak_int,
                "pipe_slot_frac_of_measured": rate * mmi * SLOTS_PER_MM / peak_int,
                "executed_pipe_slot_frac_of_measured": rate * wi["executed_mm_eq"] * SLOTS_PER_MM / peak_int,
                "hbm_gbs": rate * 8 * (pars[i].n_vars + 1) / 1e9,
            }
        e2e_bytes_per_s = e2e_value / world * (h2d + d2h) / (len(FUNCS) * ne)
        e2e_h2d_per_s = e2e_value * h2d / (len(FUNCS) * ne)          # whole job, upload direction only
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {
                "workload": "f1-f4 black-box probes (BASELINE.json configs[1]): one step = f1,f2,f3,f4 each at "
                            f"{n} points per GPU, p = 2^63-25",
                "functions": FUNCS, "prime": str(P0), "points_per_function_per_step_per_gpu": n,
                "timed_region_s": ms_total * 1e-3,
                "parallelism": f"points sharded by index over {world} GPU(s), no collective",
                "l2": "inputs larger than L2 (f1 %.0f MB, f2-f4 %.0f MB per launch); no flush needed" %
                      (20 * 8 * n / 1e6, 5 * 8 * n / 1e6),
                "launch": {"probes_per_thread": args.k or "auto", "threads_per_block": args.threads or "auto"},
            },
            "per_function": per_fn,
            "verified_bit_exact_vs_oracle": verified,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "points_per_function_per_step_per_gpu": ne, "steps": e2e_steps, "gpu_launches": int(e2e_launches),
                    "api": "ffgpu_eval (host pinned buffers, chunked H2D/kernel/D2H pipeline)",
                    "link_gb_per_s_per_gpu": e2e_bytes_per_s / 1e9,
                    "h2d_gb_per_s": e2e_h2d_per_s / 1e9,
                    "link_peak": link,
                    "frac_of_link_peak": (e2e_h2d_per_s / 1e9 / link["h2d_gbs"]) if link and link.get("h2d_gbs") else None,
                    "frac_of_both_directions_peak": ((e2e_h2d_per_s + e2e_value * d2h / (len(FUNCS) * ne)) / 1e9 / link["both_gbs_sum"])
                    if link and link.get("both_gbs_sum") else None,
                    "frac_note": "frac_of_link_peak = upload bytes/s of the whole job / pinned H2D copy rate of all its GPUs "
                                 "copying at once (87 % of the bytes of a step go up); frac_of_both_directions_peak = all bytes/s "
                                 "/ the rate with uploads and downloads running together; the larger of the two is the "
                                 "ceiling that binds (one GPU: the upload link; eight GPUs of this VM: the host's total)"},
            "e2e_line": {"value": line_value, "unit": UNIT, "h2d_bytes_per_step": int(lh2d), "d2h_bytes_per_step": int(ld2h),
                         "points_per_function_per_step_per_gpu": ne, "steps": e2e_steps, "gpu_launches": int(line_launches),
                         "api": "ffgpu_eval_line (host pinned t[], x_v = a_v*t + s_v formed on the GPU as "
                                "Reconstructor::queue_probes does on the host; same values as e2e)"},
            "e2e_ctx": e2e_ctx,
            "roofline": {
                "bound": "int_pipe", "kernel": f"eval_tape_kernel on {FUNCS[top]}",
                "achieved": imad / 1e9, "peak": peak_int / 1e9, "unit": "G IMAD thread-inst/s",
                "frac": imad / peak_int,
                "pipe_slot_frac": probes_per_s_top * mm_eq * SLOTS_PER_MM / peak_int,
                "executed": {"mm_eq_per_probe": w["executed_mm_eq"],
                             "note": "the tape shares equal sub-expressions (f1: 20 powers instead of the reference's 40); "
                                     "frac counts the reference's token list (SURVEY 8d), this block what the kernel ran",
                             "imad_frac": probes_per_s_top * w["executed_mm_eq"] * IMAD_PER_MM / peak_int,
                             "pipe_slot_frac": probes_per_s_top * w["executed_mm_eq"] * SLOTS_PER_MM / peak_int},
                "peak_source": peak_src,
                "modmul_ceiling": (lambda m: None if m is None else dict(
                    m, executed_gmm_eq_per_s=probes_per_s_top * w["executed_mm_eq"] / 1e9,
                    frac_of_mul_chain=probes_per_s_top * w["executed_mm_eq"] / 1e9 / m["mul"]))(measured_modmul_peak()),
                "algorithmic": {"mm": w["mm"], "inv": w["inv"], "ma": w["ma"], "mm_eq_per_probe": mm_eq,
                                "imad_per_mm": IMAD_PER_MM, "mm_per_inv": MM_PER_INV,
                                "pipe_slots_per_mm": SLOTS_PER_MM},
                "hbm": {"achieved": probes_per_s_top * bytes_per_probe / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": probes_per_s_top * bytes_per_probe / 1e9 / hbm_peak, "peak_source": hbm_src,
                        "bytes_per_probe": bytes_per_probe},
                "traffic": (lambda t: None if t is None else t["bytes_per_probe"] * n)(ncu_traffic(FUNCS[top])),
                "traffic_source": (lambda t: None if t is None else t["source"])(ncu_traffic(FUNCS[top])),
            },
            "configs": configs,
            "cpu_baseline": cpu,
        }
        _emit(real_stdout, line)
    if world > 1:
        dist.destroy_process_group()
    if not verified:
        raise SystemExit("verification against the oracle FAILED")


# ------------------------------------------------------------------------------------------------------
def run_e2e_ctx(ff, W, orc, n_gpus, ne, steps):
    """ONE context over all GPUs, one ffgpu_eval per function per step on a pinned host batch of n_gpus * ne points that
    the library shards in-process by point index (capi.cu eval_host: contiguous range per GPU, chunked H2D / kernel /
    D2H on every GPU's own streams, host gather by concatenation) -- what GpuBlackBox(file, vars, false, {0..N-1}) does"""
    import ctypes as C
    ctx_all = ff.Context(list(range(n_gpus)))
    n = n_gpus * ne
    ok = True
    pars, bufs, h2d, d2h = [], [], 0, 0
    for name in FUNCS:
        text, vars_ = W.benchmark(name)
        par = ff.GpuParser(ctx_all, [text], vars_)
        par.set_prime(P0)
        nv = par.n_vars
        pin = ff.lib.ffgpu_host_alloc(nv * n * 8)
        pout = ff.lib.ffgpu_host_alloc(n * 8)
        assert pin and pout, ff.last_error()
        arr = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_uint64)), shape=(nv, n))
        arr[:] = ff.synthetic_points(nv, n, P0, seed=SEED + 3)
        oarr = np.ctypeslib.as_array(C.cast(pout, C.POINTER(C.c_uint64)), shape=(1, n))
        pars.append(par)
        bufs.append((pin, arr, pout, oarr, text, vars_))
        h2d += nv * n * 8
        d2h += n * 8

    def step():
        for par, (pin, _, pout, _, _, _) in zip(pars, bufs):
            par.eval_ptr(pin, n, pout)

    step()
    step()
    l0 = ff.launch_count()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    launches = ff.launch_count() - l0
    for par, (pin, arr, pout, oarr, text, vars_) in zip(pars, bufs):
        # a sample from EVERY GPU's range
        sel = np.unique(np.concatenate([np.arange(g * ne, (g + 1) * ne, max(1, ne // 256)) for g in range(n_gpus)] + [[n - 1]]))
        want = orc.parse([text], vars_, P0).eval(np.ascontiguousarray(arr[:, sel]))
        ok = ok and bool(np.array_equal(oarr[:, sel], want))
    res = {"value": steps * len(FUNCS) * n / dt, "unit": UNIT, "n_gpus_in_one_context": n_gpus,
           "points_per_function_per_step": n, "steps": steps, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "gpu_launches": int(launches), "verified_bit_exact_vs_oracle": ok,
           "api": "one ffgpu_ctx over all GPUs, ffgpu_eval shards the host batch in-process (wall clock on rank 0)"}
    for par, (pin, _, pout, _, _, _) in zip(pars, bufs):
        par.close()
        ff.lib.ffgpu_host_free(pin)
        ff.lib.ffgpu_host_free(pout)
    ctx_all.close()
    return res


def run_cfg_factors(ff, W, orc, ctx, torch, stream, rank, world, max_over_ranks, all_ok, barrier, n=1 << 22):
    """BASELINE.json configs[2]: factors1-5, probes in primes()[0..7], points sharded over the GPUs of the job (every rank
    its contiguous range, no communication).  One handle per prime (ffgpu_tape_clone + ffgpu_set_prime: the multi-prime
    recipe -- the prime is a property of the handle, not a global as FFInt::p is); device time by CUDA events around the
    40 launches, points generated on the device outside the timed sections."""
    primes = ff.primes()[:8]
    ok = True
    per_fn = {}
    t_all = 0.0
    for name in FACTORS:
        text, vars_ = W.benchmark(name)
        base = ff.GpuParser(ctx, [text], vars_)
        base.set_prime(primes[0])
        handles = [base] + [base.clone() for _ in primes[1:]]
        for h, p in zip(handles, primes):
            h.set_prime(p)
        nv = len(vars_)
        v = torch.empty((nv, n), dtype=torch.int64, device="cuda")
        o = torch.empty((len(primes), n), dtype=torch.int64, device="cuda")
        ms = 0.0
        first = rank * n
        idx = np.arange(0, n, max(1, n // 256))
        tidx = torch.from_numpy(idx).cuda()
        for k, (h, p) in enumerate(zip(handles, primes)):
            assert ff.lib.ffgpu_fill_points(ctx.h, 0, v.data_ptr(), nv, n, n, first, n * world, 11 + k, p, stream) == 0
            h.eval_device(v.data_ptr(), n, n, o[k].data_ptr(), n, stream=stream)       # warm-up
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            h.eval_device(v.data_ptr(), n, n, o[k].data_ptr(), n, stream=stream)
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
            host = v[:, tidx].cpu().numpy().view(np.uint64)
            want = orc.parse([text], vars_, p).eval(np.ascontiguousarray(host))
            ok = ok and bool(np.array_equal(o[k:k + 1, tidx].cpu().numpy().view(np.uint64), want))
        ms = max_over_ranks(ms)
        w = base.work()
        per_fn[name] = {"probes_per_s": n * world * len(primes) / (ms * 1e-3), "ms_for_8_primes": ms,
                        "mm_eq_per_probe": w["mm"] + MM_PER_INV * w["inv"]}
        t_all += ms
        for h in handles:
            h.close()
        del v, o
    ok = all_ok(ok)
    return {"config": "factors1-5 x primes()[0..7], points sharded over the GPUs (BASELINE.json configs[2])",
            "value": len(FACTORS) * len(primes) * n * world / (t_all * 1e-3), "unit": UNIT, "n_gpus": world,
            "points_per_gpu_per_function_per_prime": n, "per_function": per_fn, "verified_bit_exact_vs_oracle": ok,
            "timing": "CUDA events around each launch, summed, max over ranks"}


def run_cfg_ibp(ff, W, orc, ctx, torch, stream, rank, world, max_over_ranks, all_ok, barrier, n_functions, tile=1 << 15, tiles=4):
    """BASELINE.json configs[3], a bounded sample: the synthetic IBP-like collection (10 000 functions, 8 variables) at
    tiles * tile points per GPU; outputs reduced on the device to per-function (sum mod p, xor) checksums; a sample of
    points checked value by value against the oracle.  (The full 1e8 points are this many tiles more of the same.)"""
    p = ff.primes()[0]
    t0 = time.perf_counter()
    funs, vars_ = W.ibp_like_collection(n_functions, 8)
    par = ff.GpuParser(ctx, funs, vars_)
    par.set_prime(p)
    compile_s = time.perf_counter() - t0
    F = par.n_out
    v = torch.empty((8, tile), dtype=torch.int64, device="cuda")
    o = torch.empty((F, tile), dtype=torch.int64, device="cuda")
    sums = torch.zeros((tiles, F, 2), dtype=torch.int64, device="cuda")
    n_total = tiles * tile * world
    ms = 0.0
    for t in range(-1, tiles):                       # t = -1: warm-up
        first = (rank * tiles + max(t, 0)) * tile
        assert ff.lib.ffgpu_fill_points(ctx.h, 0, v.data_ptr(), 8, tile, tile, first, n_total, 5, p, stream) == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        par.eval_device(v.data_ptr(), tile, tile, o.data_ptr(), tile, stream=stream)
        ff.lib.ffgpu_checksum(ctx.h, 0, o.data_ptr(), F, tile, tile, p, sums[max(t, 0)].data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        if t >= 0:
            ms += e0.elapsed_time(e1)
    # the last tile: a sample of points value by value, and the checksums of 64 functions recomputed on the host
    idx = np.arange(0, tile, max(1, tile // 16))
    host = v.cpu().numpy().view(np.uint64)
    want = orc.parse(funs, vars_, p).eval(np.ascontiguousarray(host[:, idx]))
    got_full = o[:64].cpu().numpy().view(np.uint64)
    got_sample = o[:, torch.from_numpy(idx).cuda()].cpu().numpy().view(np.uint64)
    ok = bool(np.array_equal(got_sample, want))
    sm = sums[tiles - 1].cpu().numpy().view(np.uint64)
    chk_sum = np.array([int(sum(int(x) for x in row) % p) for row in got_full], dtype=np.uint64)
    ok = ok and bool(np.array_equal(sm[:64, 0], chk_sum)) and bool(np.array_equal(sm[:64, 1], np.bitwise_xor.reduce(got_full, axis=1)))
    ok = all_ok(ok)
    ms = max_over_ranks(ms)
    w = par.work()
    x = par.tape_work()
    res = {"config": f"synthetic IBP-like collection, {F} functions, 8 variables (BASELINE.json configs[3]; bounded sample)",
           "n_gpus": world, "functions": F, "points": n_total, "tile_points": tile,
           "value": n_total / (ms * 1e-3), "unit": UNIT, "function_values_per_s": n_total * F / (ms * 1e-3),
           "function_values_per_s_per_gpu": n_total * F / (ms * 1e-3) / world, "ms": ms,
           "tape_instructions": par.n_instr, "slots": par.n_slots, "mm_eq_per_probe": w["mm"] + MM_PER_INV * w["inv"],
           "executed_mm_eq_per_probe": x["mm"] + MM_PER_INV * x["inv"],
           "gmm_eq_per_s_per_gpu": n_total / (ms * 1e-3) * (w["mm"] + MM_PER_INV * w["inv"]) / 1e9 / world,
           "compile_and_set_prime_s": compile_s, "verified_bit_exact_vs_oracle": ok,
           "shared_rows": par.n_table_rows,
           "timing": "CUDA events around evaluate + checksum kernels of every tile, summed, max over ranks"}
    # A/B: the same collection without shared rows (every power and common factor recomputed where it is used)
    off = ff.GpuParser(ctx, funs, vars_, table=False)
    off.set_prime(p)
    ms_off = 0.0
    for t in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        off.eval_device(v.data_ptr(), tile, tile, o.data_ptr(), tile, stream=stream)
        e1.record()
        torch.cuda.synchronize()
        ms_off = e0.elapsed_time(e1)
    xo = off.tape_work()
    res["without_shared_rows"] = {"ms_per_tile": max_over_ranks(ms_off), "ms_per_tile_with": ms / tiles,
                                  "executed_mm_eq_per_probe": xo["mm"] + MM_PER_INV * xo["inv"], "slots": off.n_slots}
    off.close()
    barrier()
    if rank == 0:
        # the plug-in view: ONE call of the black box with a bunch of 128 points in host memory (what a Reconstructor pool
        # thread does, utils.cpp:305-330) -- GPU through ffgpu_eval, and the reference's evaluate_pre<FFIntVec<128>> on one
        # host thread beside it (oracle/_ref when it was built, i.e. never on a box without it)
        pts = ff.synthetic_points(8, 128, p, seed=77)
        want = orc.parse(funs, vars_, p).eval(pts)
        got = par.evaluate_pre(pts)
        for _ in range(5):
            got = par.evaluate_pre(pts)
        reps = 30
        t0 = time.perf_counter()
        for _ in range(reps):
            par.evaluate_pre(pts)
        gpu_us = (time.perf_counter() - t0) / reps * 1e6
        plug = {"bunch": 128, "gpu_us_per_call": gpu_us, "gpu_function_values_per_s": 128 * F / (gpu_us * 1e-6),
                "verified_bit_exact_vs_oracle": bool(np.array_equal(got, want))}
        res["verified_bit_exact_vs_oracle"] = res["verified_bit_exact_vs_oracle"] and plug["verified_bit_exact_vs_oracle"]
        try:
            from refs import Ref, have_ref
            if have_ref():
                ref = Ref()
                ref.set_prime(p)
                rp = ref.parse(funs, vars_)
                rp.eval(pts, bunch=128, threads=1)
                t0 = time.perf_counter()
                for _ in range(3):
                    rp.eval(pts, bunch=128, threads=1)
                cpu_us = (time.perf_counter() - t0) / 3 * 1e6
                plug["reference_cpu_us_per_call_1_thread"] = cpu_us
                plug["speedup_per_call"] = cpu_us / gpu_us
                rp.close()
        except Exception as e:
            plug["reference_cpu_us_per_call_1_thread"] = repr(e)
        res["plugin_call_bunch_128"] = plug
    par.close()
    return res


def run_cfg_sweep(ff, W, orc, ctx):
    """BASELINE.json configs[1], the latency view: bunch sizes 1..4096 through ffgpu_eval with pageable host buffers,
    one caller (what ONE Reconstructor pool thread sees), and 16 concurrent callers with bunches of 128 (what the pool
    does: Reconstructor.hpp:2981-2983) with the aggregator on and off."""
    p = ff.primes()[0]
    out, ok = {}, True
    for name in FUNCS:
        text, vars_ = W.benchmark(name)
        par = ff.GpuParser(ctx, [text], vars_)
        par.set_prime(p)
        op = orc.parse([text], vars_, p)
        row = {}
        for b in [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096]:
            pts = ff.synthetic_points(par.n_vars, b, p, seed=b)
            for _ in range(10):
                got = par.evaluate_pre(pts)
            ok = ok and bool(np.array_equal(got, op.eval(pts)))
            reps = 100
            t0 = time.perf_counter()
            for _ in range(reps):
                par.evaluate_pre(pts)
            dt = (time.perf_counter() - t0) / reps
            row[str(b)] = {"us_per_call": dt * 1e6, "probes_per_s": b / dt}
        out[name] = row
        if name == "f4":
            conc = {}
            for mode in (1, 0):
                par.set_coalesce(mode)
                n_thr, reps, b = 16, 200, 128
                batches = [ff.synthetic_points(par.n_vars, b, p, seed=1000 + t) for t in range(n_thr)]
                wants = [op.eval(x) for x in batches]
                bad = []

                def worker(t):
                    for _ in range(reps):
                        g = par.evaluate_pre(batches[t], thread_id=t)
                    if not np.array_equal(g, wants[t]):
                        bad.append(t)

                l0 = ff.launch_count()
                ths = [threading.Thread(target=worker, args=(t,)) for t in range(n_thr)]
                t0 = time.perf_counter()
                [t.start() for t in ths]
                [t.join() for t in ths]
                dt = time.perf_counter() - t0
                ok = ok and not bad
                conc["aggregator_on" if mode else "aggregator_off"] = {
                    "calls_per_s": n_thr * reps / dt, "probes_per_s": n_thr * reps * b / dt,
                    "kernel_launches": int(ff.launch_count() - l0), "calls": n_thr * reps}
            par.set_coalesce(-1)
            out["f4_16_concurrent_callers_x128"] = conc
        par.close()
    return {"config": "bunch sweep 1..4096 through ffgpu_eval, pageable host buffers (BASELINE.json configs[1])",
            "unit": UNIT, "sweep": out, "verified_bit_exact_vs_oracle": ok}


def run_cfg_large_function(ff, orc, ctx, n_terms=10000):
    """The plug-in view for ONE large function (a rational function with 10^4 terms in the numerator, 5 variables): one
    call of the black box with a bunch of 1 / 128 points in host memory -- what a Reconstructor pool thread does.  The
    compiler cuts the long sums into partial sums that parallel chunks of the rows' tape evaluate (include/ffgpu.h,
    shared rows level 1); `unsplit` is the same function walked by one thread block; the reference's
    evaluate_pre<FFIntVec<128>> on one host thread stands beside it where oracle/_ref was built."""
    import random
    rng = random.Random(5)
    vs = [f"x{i}" for i in range(1, 6)]
    terms = ["%d*%s" % (rng.randint(1, 10 ** 12), "*".join(f"{v}^{rng.randint(1, 30)}" for v in rng.sample(vs, 3))) for _ in range(n_terms)]
    fun = "(" + "".join(rng.choice("+-") + t for t in terms)[1:] + ")/(" + "+".join(terms[: n_terms // 4]) + ")"
    p = ff.primes()[0]
    res = {"config": f"one rational function, {n_terms} + {n_terms // 4} terms, 5 variables; ffgpu_eval with pageable host buffers",
           "unit": "us per call"}
    pts128 = ff.synthetic_points(5, 128, p, seed=9)
    want = orc.parse([fun], vs, p).eval(pts128)
    ok = True
    for split in (True, False):
        t0 = time.perf_counter()
        par = ff.GpuParser(ctx, [fun], vs, split=split)
        par.set_prime(p)
        tc = time.perf_counter() - t0
        row = {"compile_and_set_prime_s": tc, "shared_rows": par.n_table_rows, "partial_sums": par.n_split_rows,
               "main_tape_instructions": par.n_instr}
        for b in (1, 128):
            pts = np.ascontiguousarray(pts128[:, :b])
            for _ in range(3):
                got = par.evaluate_pre(pts)
            ok = ok and bool(np.array_equal(got, want[:, :b]))
            reps = 20 if split else 5
            t0 = time.perf_counter()
            for _ in range(reps):
                par.evaluate_pre(pts)
            row[f"bunch_{b}_us"] = (time.perf_counter() - t0) / reps * 1e6
        res["split" if split else "unsplit"] = row
        par.close()
    try:
        from refs import Ref, have_ref
        if have_ref():
            ref = Ref()
            ref.set_prime(p)
            rp = ref.parse([fun], vs)
            rp.eval(pts128, bunch=128, threads=1)
            t0 = time.perf_counter()
            for _ in range(3):
                rp.eval(pts128, bunch=128, threads=1)
            res["reference_cpu_bunch_128_us_1_thread"] = (time.perf_counter() - t0) / 3 * 1e6
            one = np.ascontiguousarray(pts128[:, :1])
            t0 = time.perf_counter()
            for _ in range(5):
                rp.eval(one)
            res["reference_cpu_bunch_1_us_1_thread"] = (time.perf_counter() - t0) / 5 * 1e6
            rp.close()
    except Exception as e:
        res["reference_cpu_bunch_128_us_1_thread"] = repr(e)
    res["verified_bit_exact_vs_oracle"] = ok
    return res


def cpu_baseline(n):
    """The reference's evaluate_pre on this host's cores, bounded sample (reported, not the target)."""
    import firefly_b200 as ff
    from firefly_b200 import workloads as W
    from refs import have_ref
    cores = os.cpu_count() or 1
    t_total = 0.0
    if have_ref():
        from refs import Ref
        ref = Ref()
        ref.set_prime(P0)
        for name in FUNCS:
            text, vars_ = W.benchmark(name)
            par = ref.parse([text], vars_)
            pts = ff.synthetic_points(len(vars_), n, P0, seed=SEED)
            par.eval(pts[:, : min(n, 1 << 14)], bunch=128, threads=cores)
            best = 1e30
            for _ in range(3):
                t0 = time.perf_counter()
                par.eval(pts, bunch=128, threads=cores)
                best = min(best, time.perf_counter() - t0)
            t_total += best
        return {"value": len(FUNCS) * n / t_total, "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": f"{n} points per function of f1-f4, reference evaluate_pre<FFIntVec<128>> (DEFAULT arithmetic, "
                          f"g++ -O3), {cores} threads, best of 3"}
    from oracle_bind import Oracle
    orc = Oracle()
    n = min(n, 1 << 18)
    for name in FUNCS:
        text, vars_ = W.benchmark(name)
        par = orc.parse([text], vars_, P0)
        pts = ff.synthetic_points(len(vars_), n, P0, seed=SEED)
        t0 = time.perf_counter()
        par.eval(pts)
        t_total += time.perf_counter() - t0
    return {"value": len(FUNCS) * n / t_total, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n} points per function of f1-f4, scalar C port of evaluate_pre<FFInt>, 1 thread"}


if __name__ == "__main__":
    main()
