"""Leave-one-out cross-validation over a smoothing grid as ONE batched device-resident optimisation
(SURVEY.md §8 f1 + f4; reference: the CV loop of src/main.cpp:588-822).

Reference flow: for each smoothing value (sequentially, warm-started from the previous one) fit the full data,
then for each tip prune it (CV mask), re-optimise from the full-data optimum, and score
chi^2 += (c_tip - r_par d_tip)^2 / (r_par d_tip)  (src/main.cpp:737-757), folds in an OpenMP loop.
Here every (smoothing value, fold) pair is one vector of the batch: per-vector smoothing + per-vector CV mask
(tpl_batch_configure), one launch of the PL kernels per trial point for all of them.
"""
from __future__ import annotations

import time

import numpy as np

from . import batchopt
from . import engine as eng


def start_vector(flat, P, n_rates):
    """A feasible start in the full PL layout: the reference's start dates (get_feasible_start_dates) and one
    clock rate sum(c) / sum(durations) on every branch (the LF phase's job in the reference)."""
    n = len(flat["parents"])
    par = flat["parents"]
    dates = flat["start_dates"]
    dur = dates[par[1:]] - dates[1:]
    rate = float(np.sum(flat["c"][1:]) / np.sum(dur))
    x = np.empty(P)
    x[:n_rates] = rate
    x[n_rates:] = dates[flat["free"] == 1]
    return x


def loocv_chisq(flat, X, lane_tip, freeparams):
    """chi-square term of each vector (src/main.cpp:737-757). X: [P,B] optimised vectors, lane_tip[b] = pruned tip."""
    n = len(flat["parents"])
    par = flat["parents"]
    fp = np.asarray(freeparams)
    out = np.zeros(X.shape[1])
    for b, tip in enumerate(lane_tip):
        p = par[tip]
        if p == 0:
            sib = [j for j in range(n) if par[j] == 0 and j != tip][-1]      # the loop keeps the LAST match
            r = X[fp[sib], b]
        else:
            r = X[fp[p], b]
        hp = X[fp[n + p], b] if fp[n + p] != -1 else flat["start_dates"][p]
        d = hp - flat["start_dates"][tip]
        if d == 0:
            d = 2.2250738585072014e-308
        expe = r * d
        c = flat["c"][tip]
        out[b] = 0.0 if (d < 1e-100 or expe < 1e-100) else (c - expe) ** 2 / expe
    return out


def _minimize(plp, X0, lo, hi, nr, dscale, max_iter, native, history, verbose, **opt):
    """-> (X, F, info) with info keys iterations (max accepted steps), nfev (batched evaluations), converged [B]."""
    P, B = X0.shape
    if native:
        # chi-square is far more sensitive to the last digits of the optimum than the objective is (at small
        # smoothing the pruned tip's prediction hangs on weakly determined rates): converge tightly
        opt.setdefault("ftol", 1e-15)
        opt.setdefault("patience", 40)
        X, F, info = plp.optimize_batch(X0, lo, hi, mode=eng.GRAD_AD, history=history, max_steps=20 * max_iter,
                                        date_scale=dscale, use_graph=1, **opt)
        bad = np.nonzero(info["status"] == 4)[0]
        if bad.size:
            raise ValueError("start point infeasible for vectors %s" % bad.tolist()[:8])
        return X, F, {"iterations": int(info["iterations"].max()), "nfev": info["steps"],
                      "converged": (info["status"] >= 1) & (info["status"] <= 3), "device_ms": info["device_ms"],
                      "launches": info["launches"]}
    opt = batchopt.BatchedLBFGS(plp, P, B, nr, lo, hi, date_scale=dscale)
    return opt.minimize(X0, max_iter=max_iter, verbose=verbose)


def loocv(flat, smoothing_grid, tips, device=0, log_pen=False, max_iter=3000, verbose=False, native=True, history=8, **opt):
    """-> dict(chisq per smoothing value, best smoothing, timings).  tips: node numbers in fold order
    (externalNodes order, main.cpp:603-608).  native=True: the optimiser is the library's own device-resident
    L-BFGS (tpl_optimize_batch); False: the torch-vector-arithmetic prototype of batchopt.py."""
    if not native:
        import torch

        torch.cuda.set_device(device)
    n = len(flat["parents"])
    plp = eng.PLCalcParallel.from_flat(flat, device=device)
    plp.set_log_pen(log_pen)
    fp, P = eng.generate_param_order_vector(flat["free"], lf=False)
    plp.set_freeparams(P, False, fp)
    nr = n - 1
    lo, hi = batchopt.reference_bounds(flat, fp, P)
    K, T = len(smoothing_grid), len(tips)
    t0 = time.perf_counter()
    # ---- full-data fits: one vector per smoothing value (padded to an even batch) ----
    B1 = K + (K % 2)
    lam1 = np.array(list(smoothing_grid) + [smoothing_grid[-1]] * (B1 - K), dtype=np.float64)
    plp.batch_configure(B1, lane_smoothing=lam1)
    x0 = start_vector(flat, P, nr)
    par = flat["parents"]
    dscale = float(np.median(np.maximum(flat["start_dates"][par[1:]] - flat["start_dates"][1:], 1e-12)))
    Xfull, Ffull, info1 = _minimize(plp, np.repeat(x0[:, None], B1, axis=1), lo, hi, nr, dscale, max_iter, native, history, verbose, **opt)
    t1 = time.perf_counter()
    # ---- folds: one vector per (smoothing value, pruned tip) ----
    B2 = K * T
    B2p = B2 + (B2 % 2)
    lam2 = np.empty(B2p)
    masks, lane_tip = [], []
    X0 = np.empty((P, B2p))
    for k in range(K):
        for j, tip in enumerate(tips):
            b = k * T + j
            lam2[b] = smoothing_grid[k]
            masks.append([int(tip)])
            lane_tip.append(int(tip))
            X0[:, b] = Xfull[:, k]
    for b in range(B2, B2p):
        lam2[b] = lam2[B2 - 1]; masks.append(masks[B2 - 1]); lane_tip.append(lane_tip[B2 - 1]); X0[:, b] = X0[:, B2 - 1]
    plp.batch_configure(B2p, lane_smoothing=lam2, masks=masks)
    Xcv, Fcv, info2 = _minimize(plp, X0, lo, hi, nr, dscale, max_iter, native, history, verbose, **opt)
    t2 = time.perf_counter()
    terms = loocv_chisq(flat, Xcv, lane_tip, fp)[:B2].reshape(K, T)
    chisq = terms.sum(axis=1)                      # fold order, fixed
    plp.close()
    return {"smoothing": list(smoothing_grid), "chisq": chisq.tolist(), "best": float(smoothing_grid[int(np.argmin(chisq))]),
            "full_objective": Ffull[:K].tolist(), "full_seconds": t1 - t0, "cv_seconds": t2 - t1,
            "full_info": {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in info1.items()},
            "cv_iterations": info2["iterations"], "cv_nfev": info2["nfev"], "cv_converged": int(info2["converged"][:B2].sum()),
            "full_device_ms": info1.get("device_ms"), "cv_device_ms": info2.get("device_ms"),
            "launches": (info1.get("launches") or 0) + (info2.get("launches") or 0),
            "n_vectors": B2, "terms": terms}
