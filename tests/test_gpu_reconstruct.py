"""Identical reconstructions (BASELINE.json configs[0], configs[4]): the unmodified reference
Reconstructor run over the reference CPU black box and over firefly::GpuBlackBox must return the same
rational functions.  Uses oracle/_ref/gpu_reconstruct (built from oracle/gpu_reconstruct.cpp where the
reference sources are available; the binary travels to the GPU box)."""
import os
import subprocess

import numpy as np
import pytest

import firefly_b200 as ff
from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "oracle", "_ref", "gpu_reconstruct")


CASES = [
    ("s_y_4_v", 1, 4, "shift"),       # configs[0]: s_y_4_v.m, bunch_size = 1 (test/test_normal.cpp:87-91)
    ("s_y_4_v", 4, 4, "shift"),       # the reference's test_bunch.cpp:67-71
    ("s_y_1_v", 128, 2, "shift"),
    ("f1", 64, 4, "shift"),
    ("factors3", 16, 4, "shift"),
    # the rest of the reference's own test matrix for this path: one variable without scans (test_normal.cpp:79-83)
    # and in bunches (test_bunch.cpp:74-78), the safe interpolation mode on s_y_safe.m, s_y_1_v.m and s_y_4_v.m
    # (test_safe.cpp:77-96, test_bunch.cpp:81-85)
    ("s_y_1_v", 1, 4, "none"),
    ("s_y_1_v", 4, 4, "shift"),
    ("s_y_safe", 1, 4, "safe"),
    ("s_y_safe", 4, 4, "safe"),
    ("s_y_1_v", 1, 4, "safe"),
    ("s_y_4_v", 4, 4, "safe"),
]
# configs[4]: the full f4 reconstruction (106 668 probes, about 2.5 minutes for the CPU and the GPU run together --
# the interpolation is sequential host work).  Runs wherever the box has the cores the reference's thread pool is
# given here; FFGPU_SKIP_SLOW_TESTS=1 leaves it out.
if (os.cpu_count() or 1) >= 8 and not os.environ.get("FFGPU_SKIP_SLOW_TESTS"):
    CASES.append(("f4", 128, 16, "shift"))


def _case_id(c):
    name, bunch, threads, mode = c
    return f"{name}-{bunch}-{threads}" + ("" if mode == "shift" else f"-{mode}")


@pytest.mark.parametrize("name,bunch,threads,mode", CASES, ids=[_case_id(c) for c in CASES])
def test_reconstruction_identical(tmp_path, name, bunch, threads, mode):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/gpu_reconstruct not built (needs /root/reference at build time)")
    g = load_golden(name)
    f = tmp_path / (name + ".m")
    f.write_text(";\n".join(g["functions"]) + ";\n")
    # the full f4 run also exercises the aggregator of concurrent small calls under the real thread pool (off by default)
    env = dict(os.environ, FFGPU_RECONSTRUCT_MODE=mode)
    if name == "f4":
        env["FFGPU_COALESCE"] = "1"
    r = subprocess.run([BIN, str(f), str(bunch), str(threads), *g["vars"]], cwd=tmp_path, capture_output=True,
                       text=True, timeout=1200, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "0 differ" in r.stdout
    # the GPU black box really ran on the GPU
    launches = int(r.stdout.strip().split()[-1])
    assert launches > 0
    # the reference's own "average time of the black-box probe" for both black boxes, kept for the record
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("black box")]
    assert len(lines) == 2, r.stdout[-2000:]
    print("\n".join([f"[{name} bunch {bunch} threads {threads}] " + ln for ln in lines]))
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "reconstruct_black_box_times.log"), "a") as log:
        for ln in lines:
            log.write(f"{name} bunch={bunch} threads={threads} mode={mode} launches={launches} {ln}\n")
    if name == "f4":
        # 16 pool threads with bunches of up to 128 points: concurrent calls share launches (ffgpu_set_coalesce);
        # round 1 needed 28 767 launches for this run
        assert launches * 5 <= 28767, f"{launches} launches: the aggregator did not coalesce the pool's calls"


def test_reference_recorded_probes(ctx, tmp_path):
    """The probes the reference ships (examples/firefly_executable/probes/*.gz, three primes) come out of the offline
    service line for line, and out of ffgpu_eval_line from (anchor points, shifts, zi_order, t) alone."""
    import gzip
    from firefly_b200 import offline
    g = load_golden("firefly_executable")
    fn = tmp_path / "funs.m"
    fn.write_text(";\n".join(g["functions"]) + ";\n")
    par = ff.GpuParser(ctx, g["functions"], g["vars"])
    for k, case in enumerate(g["cases"]):
        p = case["prime"]
        assert p == ff.primes()[k]
        req = tmp_path / f"requested_{k}.gz"
        with gzip.open(req, "wt") as f:
            for j, pr in enumerate(case["probes"]):
                f.write(" ".join(str(z) for z in pr["zi_order"]) + f" | {pr['t']} | "
                        + " ".join(str(int(x)) for x in case["points"][:, j]) + "\n")
        out = tmp_path / "probes" / f"{k}.gz"
        assert offline.serve(str(fn), g["vars"], k, str(req), str(out)) == len(case["probes"])
        with gzip.open(out, "rt") as f:
            assert f.read().splitlines() == [pr["line"] for pr in case["probes"]]
        # the same probes from the line parameters: x_1 = t + s_1, x_i = anchor_{i-1}^zi_{i-1} * t + s_i
        par.set_prime(p)
        anchors = [int(x) for x in g["anchor_points"][k]]
        shifts = np.array([int(x) for x in g["shifts"][k]], dtype=np.uint64)
        for j, pr in enumerate(case["probes"]):
            a = np.array([1] + [pow(anchors[i], pr["zi_order"][i], p) for i in range(2)], dtype=np.uint64)
            t = np.array([int(pr["t"])], dtype=np.uint64)
            assert np.array_equal(par.evaluate_line(a, shifts, t)[:, 0], case["values"][:, j]), (k, j)
    par.close()


def test_added_trimmed_and_cloned_collections_on_device(ctx, oracle):
    """ffgpu_add_function / ffgpu_trim / ffgpu_tape_clone with a GPU behind them: a stale tape refuses to evaluate,
    set_prime rebuilds and uploads it (new device buffers), small and chunked calls give the oracle's values."""
    g = load_golden("s_y_4_v")
    funs, vars_ = g["functions"], g["vars"]
    p = ff.primes()[1]
    pts = ff.synthetic_points(4, 300, p, seed=3)
    pts[:, 0] = 0
    big = ff.synthetic_points(4, 70001, p, seed=4)

    def want(fs, x):
        return oracle.parse(fs, vars_, p).eval(x)

    par = ff.GpuParser(ctx, funs[:10], vars_)
    par.set_prime(p)
    assert np.array_equal(par.evaluate_pre(pts), want(funs[:10], pts))
    for f in funs[10:]:
        par.add_otf(f)
    with pytest.raises(ff.FFGpuError, match="set_prime"):
        par.evaluate_pre(pts)
    par.set_prime(p)
    assert par.n_out == len(funs)
    assert np.array_equal(par.evaluate_pre(pts), want(funs, pts))
    twin = par.clone()
    keep = [0, 3, 4, 11, 18]
    assert par.trim(keep) == {k: i for i, k in enumerate(keep)}
    par.set_prime(p)
    kept = [funs[k] for k in keep]
    assert np.array_equal(par.evaluate_pre(pts), want(kept, pts))
    par.set_chunk(1 << 14)                              # several pipelined chunks on the rebuilt tape
    assert np.array_equal(par.evaluate_pre(big), want(kept, big))
    # the clone kept the untrimmed collection, and outlives the original
    par.close()
    assert np.array_equal(twin.evaluate_pre(pts), want(funs, pts))
    p2 = ff.primes()[6]
    twin.set_prime(p2)
    pts2 = ff.synthetic_points(4, 129, p2, seed=5)
    assert np.array_equal(twin.evaluate_pre(pts2), oracle.parse(funs, vars_, p2).eval(pts2))
    twin.close()


PARSER_BIN = os.path.join(ROOT, "oracle", "_ref", "parser_dropin")


@pytest.mark.parametrize("name", ["s_y_4_v", "f2", "factors3"])
def test_parser_mirror_matches_reference(tmp_path, name):
    """firefly::GpuShuntingYardParser, the class a user black box holds instead of ShuntingYardParser: the values of
    evaluate_pre<FFInt / FFIntVec<4> / FFIntVec<128>> over two primes, on copies, after add_otf, trim,
    add_otf_precompute and parse_function, and (s_y_4_v) the reference's test_bunch black box reconstructed over it,
    uninterrupted and resumed from states a CPU run saved (test/test_save_write.cpp's flow: the GPU black box has to
    reproduce ff_save/validation.gz, the reference's own known-answer check, Reconstructor.hpp:626-657)."""
    if not os.path.exists(PARSER_BIN):
        pytest.skip("oracle/_ref/parser_dropin not built (needs /root/reference at build time)")
    g = load_golden(name)
    f = tmp_path / (name + ".m")
    f.write_text(";\n".join(g["functions"]) + ";\n")
    r = subprocess.run([PARSER_BIN, "gpu", str(f), *g["vars"]], cwd=tmp_path, capture_output=True, text=True,
                       timeout=1200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "0 differ" in r.stdout
    assert int(r.stdout.strip().split()[-1]) > 0
