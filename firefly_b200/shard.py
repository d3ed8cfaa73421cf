"""Point-index sharding of a probe batch over ranks / GPUs.

Probes are independent, so the only multi-GPU logic is index arithmetic: rank r of W owns a contiguous
range of the global point index space, evaluates it on its own GPU and the pieces are concatenated along
the point axis of [function][point].  The same rule is used inside libffgpu for the GPUs of one context
(capi.cu, ffgpu_eval) and by bench.py for one-process-per-GPU runs.
"""
from __future__ import annotations

import numpy as np


def rank_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) of rank `rank`: ceil(total/world) points per rank, the tail ranks may be short
    or empty (the rule of ffgpu_eval)."""
    per = (total + world - 1) // world
    lo = min(total, rank * per)
    return lo, min(total, lo + per)


def concat_by_rank(pieces) -> np.ndarray:
    """[function][point] pieces in rank order -> the full result."""
    return np.concatenate([np.asarray(p) for p in pieces], axis=1)
