import json, sys, time
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from golden_util import Golden
from treepl_b200 import batchopt, cv, engine as eng
lams = list(10.0 ** np.linspace(4, -3, 16))
for name in ("clock",):
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
    for max_cg in (400, 1000):
        method = 1
        X, F, info = plp.optimize_batch(np.repeat(x0[:, None], B, axis=1), lo, hi, method=method, date_scale=ds, max_steps=40000, use_graph=1,
                                        max_cg=max_cg)
        f, g = plp.calc_pl_grad_batch(X, mode=eng.GRAD_AD)
        gz = g.copy(); gz[: n - 1] *= X[: n - 1]; gz[n - 1:] *= ds
        print("max_cg", max_cg, "steps", info["steps"], "mean_log10_step", info["mean_log10_step"], "cg", info["cg_iterations"])
        for b in range(B):
            dur = []
            print("  lam %9.3g status %d iters %5d F %.10g |g|max %.3e  min rate %.3e max rate %.3e" % (lams[b], info["status"][b], info["iterations"][b], F[b], np.abs(gz[:, b]).max(), X[: n - 1, b].min(), X[: n - 1, b].max()))
    dates = flat["start_dates"]; d0 = dates[par[1:]] - dates[1:]
    print("durations start: min %.3e median %.3e; c min %g; nbar %d" % (d0.min(), np.median(d0), flat["c"][1:].min(), int(((flat["pen_min"] != -1) | (flat["pen_max"] != -1)).sum())))
