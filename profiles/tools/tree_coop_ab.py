import sys, os, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from golden_util import Golden
from treepl_b200 import cv, engine as eng, batchopt, treeflat as tf

def tn100k(B, coop, steps=60):
    flat = bench.build_problem(100000)
    n = len(flat["parents"])
    fp, P = tf.generate_param_order_vector(flat["free"], lf=False)
    plp = eng.PLCalcParallel.from_flat(flat, device=0)
    plp.set_freeparams(P, False, fp)
    lam = 10.0 ** np.linspace(3, -1, B)
    plp.batch_configure(B, lane_smoothing=lam)
    lo, hi = batchopt.reference_bounds(flat, fp, P)
    x0 = cv.start_vector(flat, P, n - 1)
    par = flat["parents"]
    dscale = float(np.median(np.maximum(flat["start_dates"][par[1:]] - flat["start_dates"][1:], 1e-12)))
    X, F, info = plp.optimize_batch(np.repeat(x0[:, None], B, axis=1), lo, hi, mode=eng.GRAD_AD, date_scale=dscale, use_graph=1, method=1, check_every=10, max_steps=steps, tree_coop=coop)
    plp.close()
    return {"job": "tn100k", "B": B, "coop": coop, "steps": info["steps"], "device_ms": info["device_ms"], "launches": info["launches"], "Fsum": float(F.sum())}

def restarts(B, coop):
    flat = dict(Golden("big_test").flat)
    n = len(flat["parents"])
    plp = eng.PLCalcParallel.from_flat(flat, device=0)
    fp, P = eng.generate_param_order_vector(flat["free"], lf=False)
    plp.set_freeparams(P, False, fp)
    plp.smoothing = 1.0
    nr = n - 1
    lo, hi = batchopt.reference_bounds(flat, fp, P)
    x0 = cv.start_vector(flat, P, nr)
    X0 = np.repeat(x0[:, None], B, axis=1)
    for b in range(B):
        X0[:nr, b] *= np.exp(0.3 * np.random.RandomState(1000003 + b).randn(nr))
    plp.batch_configure(B)
    par = flat["parents"]
    dscale = float(np.median(np.maximum(flat["start_dates"][par[1:]] - flat["start_dates"][1:], 1e-12)))
    t1 = time.perf_counter()
    X, F, info = plp.optimize_batch(X0, lo, hi, mode=eng.GRAD_AD, history=8, date_scale=dscale, use_graph=1, method=1, check_every=50, max_steps=60000, tree_coop=coop)
    t2 = time.perf_counter()
    plp.close()
    return {"job": "restarts", "B": B, "coop": coop, "opt_s": t2 - t1, "device_ms": info["device_ms"], "steps": info["steps"], "launches": info["launches"], "Fmin": float(F.min()), "Fsum": float(F.sum())}

def loocv_big(coop):
    flat = bench.load_big_test()
    tips = [int(t) for t in np.flatnonzero(flat["child_counts"] == 0)]
    t0 = time.perf_counter()
    res = cv.loocv(dict(flat), bench.BIG_GRID, tips, device=0, max_iter=6000, tree_coop=coop)
    dt = time.perf_counter() - t0
    return {"job": "loocv_big", "coop": coop, "wall_s": dt, "full_s": res["full_seconds"], "folds_s": res["cv_seconds"], "full_dev_ms": res["full_device_ms"], "folds_dev_ms": res["cv_device_ms"], "chisq": res["chisq"][:3]}

tn100k(2, 0, 10)   # warm
for rep in range(2):
    for coop in (0, -1):
        for B in (2, 16, 64, 256):
            print(json.dumps(tn100k(B, coop)), flush=True)
        for B in (32, 128, 1024, 4096):
            print(json.dumps(restarts(B, coop)), flush=True)
        print(json.dumps(loocv_big(coop)), flush=True)
