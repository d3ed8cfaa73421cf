#!/usr/bin/env python
"""Development tool: the reference's average black-box time for its CPU black box and for GpuBlackBox (the harness
oracle/_ref/gpu_reconstruct), with the aggregator of concurrent small calls on and off (FFGPU_COALESCE).
    python scripts/reconstruct_times.py [case ...]      cases: s_y_4_v:1:4 f1:64:4 factors3:16:4 f4:128:16"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "gpu_reconstruct")
cases = sys.argv[1:] or ["s_y_4_v:1:4", "s_y_4_v:4:4", "f1:64:4", "factors3:16:4"]
for case in cases:
    name, bunch, threads = case.split(":")
    with open(os.path.join(ROOT, "tests", "golden", name + ".json")) as f:
        g = json.load(f)
    with tempfile.TemporaryDirectory() as tmp:
        fn = os.path.join(tmp, name + ".m")
        with open(fn, "w") as f:
            f.write(";\n".join(g["functions"]) + ";\n")
        for mode in ("1", "0"):
            env = dict(os.environ, FFGPU_COALESCE=mode)
            r = subprocess.run([BIN, fn, bunch, threads, *g["vars"]], cwd=tmp, capture_output=True, text=True, env=env, timeout=1500)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith("black box")]
            ok = "0 differ" in r.stdout
            for ln in lines:
                print(f"{name} bunch={bunch} threads={threads} aggregator={mode} identical={ok} {ln}", flush=True)
