"""Bounds of the batched optimisers: exactly those the reference adapters build (optimize_tnc.cpp:116-147)."""
from __future__ import annotations

import numpy as np

LARGE = 1e15


def reference_bounds(flat, freeparams, P):
    """Bounds exactly as the reference adapters build them (optimize_tnc.cpp:116-147)."""
    n = len(flat["parents"])
    lo = np.zeros(P)
    hi = np.full(P, LARGE)
    fp = np.asarray(freeparams)
    for i in range(n):
        if fp[i] != -1:
            lo[fp[i]] = 1e-40
            hi[fp[i]] = LARGE
    for i in range(n):
        s = fp[n + i]
        if s != -1:
            lo[s] = flat["min"][i]
            hi[s] = flat["max"][i]
    lo[lo == 0] = 1e-40
    return lo, hi
