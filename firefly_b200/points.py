"""Synthetic probe points -- pure numpy, no library behind it (bench.py's reference arm imports this file alone).

Host twin of the device generator `ffgpu_fill_points` (firefly_b200/csrc/kernels.cu, fill_points_kernel):
    x[v][j] = 1 + splitmix64(seed + v * total_points + first_point + j) mod (p - 1)          in [1, p - 1]
a counter-based generator, so any shard of the global point index range can be produced independently
(the reference draws its points with xoshiro / pcg seeded by the time, BaseReconst.cpp:140-152, 218-273).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

DEFAULT_SEED = 0x4D595DF4D0F33173


def synthetic_points(n_vars: int, n_points: int, p: int, seed: int = DEFAULT_SEED, first_point: int = 0,
                     total_points: Optional[int] = None) -> np.ndarray:
    """[n_vars][n_points] uint64, identical to what ffgpu_fill_points writes for the same arguments."""
    total = n_points if total_points is None else total_points
    with np.errstate(over="ignore"):
        v = np.arange(n_vars, dtype=np.uint64).reshape(-1, 1)
        j = np.arange(n_points, dtype=np.uint64).reshape(1, -1)
        z = np.uint64(seed) + v * np.uint64(total) + np.uint64(first_point) + j
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        return np.uint64(1) + z % np.uint64(p - 1)
