#!/usr/bin/env python
"""Robustness sweep of the batched truncated Newton optimiser: every example tree of the reference x 16
smoothing values 10^4 ... 10^-3, clock-like start, one vector per smoothing value; prints one JSON object per tree with the
status histogram (index = status code: 0 step cap, 1 gradient converged, 2 objective converged, 3 step/stall, 4 infeasible
start, 5 non-finite), steps, wall seconds, Newton iterations and the largest relative projected gradient.
    python profiles/tools/tn_robustness_sweep.py [copies] > profiles/r02_tn_sweep.json          (needs a B200)
copies (default 1): every smoothing value that many times - 2 makes a 32-vector batch, which takes the software-pipelined
Hessian-vector kernel and the cooperative level sweeps instead of the small-batch forms."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import Golden                                     # noqa: E402
from treepl_b200 import batchopt, cv, engine as eng                # noqa: E402

out = {}
for name in ("M1155", "big_test", "clock", "pl4203", "test", "test_two_cal", "vib"):
    flat = dict(Golden(name).flat)
    n = len(flat["parents"])
    fp, P = eng.generate_param_order_vector(flat["free"], lf=False)
    lams = np.repeat(10.0 ** np.linspace(4, -3, 16), int(sys.argv[1]) if len(sys.argv) > 1 else 1)
    for log_pen in (False, True):
        plp = eng.PLCalcParallel.from_flat(flat, device=0)
        plp.set_log_pen(log_pen)
        plp.set_freeparams(P, False, fp)
        plp.batch_configure(len(lams), lane_smoothing=lams)
        lo, hi = batchopt.reference_bounds(flat, fp, P)
        x0 = cv.start_vector(flat, P, n - 1)
        par = flat["parents"]
        dscale = float(np.median(np.maximum(flat["start_dates"][par[1:]] - flat["start_dates"][1:], 1e-12)))
        t = time.perf_counter()
        X, F, info = plp.optimize_batch(np.repeat(x0[:, None], len(lams), axis=1), lo, hi, method=1, date_scale=dscale, use_graph=1,
                                        max_steps=40000)
        wall = time.perf_counter() - t
        f, g = plp.calc_pl_grad_batch(X, mode=eng.GRAD_AD)
        # projected gradient in the optimiser's variables (ln r, h / scale), relative to |f|
        gz = g.copy()
        gz[: n - 1] *= X[: n - 1]
        gz[n - 1:] *= dscale
        onlo = X <= lo[:, None] * (1 + 1e-9)
        onhi = X >= hi[:, None] * (1 - 1e-9)
        gz[(onlo & (gz > 0)) | (onhi & (gz < 0))] = 0.0
        rel = np.abs(gz).max(axis=0) / np.maximum(np.abs(f), 1.0)
        st = np.bincount(info["status"], minlength=6).tolist()
        out["%s%s" % (name, " log_pen" if log_pen else "")] = {
            "tree": name, "log_pen": log_pen, "nodes": n, "status": st, "at_cap": st[0], "steps": int(info["steps"]), "wall_s": wall,
            "newton_iters": [int(info["iterations"].min()), int(info["iterations"].max())], "cg": int(info["cg_iterations"]),
            "max_rel_grad": float(rel.max()), "objective": [float(v) for v in F]}
        plp.close()
print(json.dumps(out))
