"""Offline probe exchange with the stock `firefly` executable (reference: firefly.cpp:137-169).

FireFly can run without a black box: it writes the probes it needs to `requested_probes.gz`
(Reconstructor::write_requested_probes_to_file, Reconstructor.hpp:3821-3846), one line per probe

    zi_1 ... zi_{n-1} | t | x_1 ... x_n

and reads the answers for prime number k from `probes/k.gz`
(Reconstructor::load_precomputed_probes_from_file, Reconstructor.hpp:3751-3818), one line per probe

    zi_1 ... zi_{n-1} | t | f_1 ... f_m

This module is the GPU evaluator in between: parse the request, evaluate all points in one batch through
the C ABI (GpuParser.evaluate_pre -> ffgpu_eval), write the answer file.  Only text handling happens here.

    python -m firefly_b200.offline functions.m x1,x2,x3 --prime-index 0 [--requested requested_probes.gz] [--out probes/0.gz]
"""
from __future__ import annotations

import argparse
import gzip
import os
from typing import List, Sequence, Tuple

import numpy as np


def read_requested(path: str, n_vars: int) -> Tuple[List[str], List[int], np.ndarray]:
    """-> (zi_order strings, t values, points [n_vars][n_probes]) of a requested_probes file"""
    orders, ts, cols = [], [], []
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as f:
        for ln, line in enumerate(f, 1):
            line = line.rstrip("\n")
            if not line.strip():
                continue
            parts = line.split("|")
            if len(parts) != 3:
                raise ValueError(f"{path}:{ln}: expected 'zi_order | t | values'")
            vals = [int(x) for x in parts[2].split()]
            if len(vals) != n_vars:
                raise ValueError(f"{path}:{ln}: {len(vals)} coordinates, the collection has {n_vars} variables")
            orders.append(" ".join(parts[0].split()))
            ts.append(int(parts[1].strip()))
            cols.append(vals)
    pts = np.array(cols, dtype=np.uint64).T.reshape(n_vars, len(cols)) if cols else np.zeros((n_vars, 0), dtype=np.uint64)
    return orders, ts, np.ascontiguousarray(pts)


def write_probes(path: str, orders: Sequence[str], ts: Sequence[int], values: np.ndarray) -> None:
    """values: [n_out][n_probes] -> a probes/<prime>.gz file"""
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "wt") as f:
        for j, (o, t) in enumerate(zip(orders, ts)):
            f.write(f"{o} | {t} | " + " ".join(str(int(v)) for v in values[:, j]) + "\n")


def read_probes(path: str) -> Tuple[List[str], List[int], np.ndarray]:
    orders, ts, cols = [], [], []
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as f:
        for line in f:
            if not line.strip():
                continue
            a, t, v = line.split("|")
            orders.append(" ".join(a.split()))
            ts.append(int(t.strip()))
            cols.append([int(x) for x in v.split()])
    return orders, ts, np.array(cols, dtype=np.uint64).T if cols else np.zeros((0, 0), dtype=np.uint64)


def serve(functions_file: str, variables: Sequence[str], prime_index: int, requested: str = "requested_probes.gz",
          out: str | None = None, devices: Sequence[int] | None = None) -> int:
    """Evaluate every requested probe modulo primes()[prime_index] on the GPU(s); returns the probe count."""
    import firefly_b200 as ff
    ctx = ff.Context(devices)
    par = ff.GpuParser(ctx, file=functions_file, vars=list(variables))
    par.set_prime(ff.primes()[prime_index])
    orders, ts, pts = read_requested(requested, par.n_vars)
    vals = par.evaluate_pre(pts) if pts.shape[1] else np.zeros((par.n_out, 0), dtype=np.uint64)
    write_probes(out or os.path.join("probes", f"{prime_index}.gz"), orders, ts, vals)
    par.close()
    ctx.close()
    return len(ts)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("functions")
    ap.add_argument("variables", help="comma separated, in the order given to FireFly")
    ap.add_argument("--prime-index", type=int, default=0)
    ap.add_argument("--requested", default="requested_probes.gz")
    ap.add_argument("--out", default=None)
    a = ap.parse_args(argv)
    n = serve(a.functions, a.variables.split(","), a.prime_index, a.requested, a.out)
    print(f"{n} probes evaluated")


if __name__ == "__main__":
    main()
