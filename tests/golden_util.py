"""Loader for tests/golden/*.npz (written by oracle/make_golden.py from the reference itself)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FLAT_KEYS = ("parents", "child_counts", "children", "free", "c", "lf", "min", "max", "pen_min", "pen_max",
             "start_dates", "start_rates", "start_durations")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = z
        self.meta = json.loads(bytes(z["meta"]).decode())
        self.flat = {k: z["flat_" + k] for k in FLAT_KEYS}
        self.evals = self.meta["evals"]

    def arr(self, k, what):
        key = "e%d_%s" % (k, what)
        return self.z[key] if key in self.z.files else None


def rel_err(a, b):
    """max_i |a_i - b_i| / max(|b_i|, 1e-8 * ||b||_inf)  (SURVEY.md §7 'parity gates')."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if b.size == 0:
        return 0.0
    bad = ~np.isfinite(b)
    if bad.any():       # inf / NaN entries of the reference must be reproduced as such
        same = (a[bad] == b[bad]) | (np.isnan(a[bad]) & np.isnan(b[bad]))
        if not same.all():
            return float("inf")
        a, b = a[~bad], b[~bad]
        if b.size == 0:
            return 0.0
    s = float(np.max(np.abs(b)))
    if s == 0.0:
        return float(np.max(np.abs(a)))
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-8 * s)))
