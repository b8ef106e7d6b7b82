"""Robustness sweep of the truncated-Newton optimiser over all example trees and a wide smoothing range (scratch)."""
import json, sys, time
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from golden_util import Golden, golden_names
from treepl_b200 import batchopt, cv, engine as eng
out = {}
lams = list(10.0 ** np.linspace(4, -3, 16))
for name in golden_names():
    flat = dict(Golden(name).flat)
    n = len(flat["parents"])
    plp = eng.PLCalcParallel.from_flat(flat, device=0)
    fp, P = eng.generate_param_order_vector(flat["free"], lf=False)
    plp.set_freeparams(P, False, fp)
    B = len(lams)
    plp.batch_configure(B, lane_smoothing=np.array(lams))
    lo, hi = batchopt.reference_bounds(flat, fp, P)
    x0 = cv.start_vector(flat, P, n - 1)
    par = flat["parents"]
    ds = float(np.median(np.maximum(flat["start_dates"][par[1:]] - flat["start_dates"][1:], 1e-12)))
    t = time.perf_counter()
    X, F, info = plp.optimize_batch(np.repeat(x0[:, None], B, axis=1), lo, hi, method=1, date_scale=ds, max_steps=40000, use_graph=1)
    wall = time.perf_counter() - t
    f, g = plp.calc_pl_grad_batch(X, mode=eng.GRAD_AD)
    gz = g.copy(); gz[: n - 1] *= X[: n - 1]
    r = {"tree": name, "nodes": n, "status": np.bincount(info["status"], minlength=6).tolist(), "steps": info["steps"], "wall_s": wall,
         "newton_iters": [int(info["iterations"].min()), int(info["iterations"].max())], "cg": info["cg_iterations"],
         "max_rel_grad": float(np.max(np.abs(gz).max(axis=0) / np.abs(f)))}
    out[name] = r
    print(json.dumps(r), flush=True)
    plp.close()
json.dump(out, open("gpurun_out/tn_sweep.json", "w"), indent=1)
