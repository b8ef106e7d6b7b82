"""End-to-end dating of one tree on the GPU: what `treePL <config>.cppr8s` does after its set-up phase
(reference src/main.cpp:497-880), on top of the batched device-resident optimiser:
    [cross-validation over the smoothing grid -> selected smoothing] -> final PL fit -> dated tree, rate tree, cv.out.
The reference's Langley-Fitch pre-fit, annealing rounds and `prime` / `thorough` loops are not reproduced: the
truncated-Newton fit from the clock-like start reaches the same optima."""
from __future__ import annotations

import numpy as np

from . import batchopt, config, cv, output
from . import engine as eng


def fit(flat, smoothing, device=0, log_pen=False, anneal=False, **opt):
    """One PL fit at `smoothing` -> (x [P], objective, freeparams, converged).  anneal=True runs the reference's
    annealing phase (optimize_full, utils.cpp:683-689) on the start vector first; the optimum does not depend on it."""
    n = len(flat["parents"])
    plp = eng.PLCalcParallel.from_flat(flat, device=device)
    plp.set_log_pen(log_pen)
    fp, P = eng.generate_param_order_vector(flat["free"], lf=False)
    plp.set_freeparams(P, False, fp)
    lo, hi = batchopt.reference_bounds(flat, fp, P)
    plp.batch_configure(2, lane_smoothing=np.array([smoothing, smoothing], dtype=np.float64))
    x0 = cv.start_vector(flat, P, n - 1)
    par = flat["parents"]
    dscale = float(np.median(np.maximum(flat["start_dates"][par[1:]] - flat["start_dates"][1:], 1e-12)))
    X0 = np.repeat(x0[:, None], 2, axis=1)
    if anneal:
        X0, _, _, _ = plp.anneal_batch(X0)
    X, F, info = cv._minimize(plp, X0, lo, hi, n - 1, dscale, 3000, 8, False, **opt)
    plp.close()
    return X[:, 0], float(F[0]), fp, bool(info["converged"][0])


def date_tree(flat, names, numsites, smooth=10.0, do_cv=False, randomcv=False, cvstart=1000.0, cvstop=0.1, cvmultstep=0.1,
              randomcviter=10, randomcvsamp=0.1, outfile=None, cvoutfile="cv.out", device=0, log_pen=False, seed=1, **opt):
    """-> dict(smoothing, objective, dated_tree, rate_tree, cv).  Option names and defaults are the reference's control-file
    keys (src/main.cpp:62, 71-73, 100-270)."""
    table = None
    if do_cv:
        grid = config.smoothing_grid(cvstart, cvstop, cvmultstep)     # main.cpp:275-284,668,814
        tips = np.flatnonzero(np.asarray(flat["child_counts"]) == 0)
        if randomcv:
            groups = cv.sample_random_groups(flat, tips, randomcviter, randomcvsamp, seed=seed)
        else:
            # LOOCV fold order = externalNodes order (main.cpp:603-608); chi-square sums do not depend on it beyond rounding
            groups = [[int(t)] for t in tips]
        res = cv.cross_validate(flat, grid, groups, randomcv=randomcv, device=device, log_pen=log_pen, **opt)
        # first strict minimum in grid order, as `if(chisq < lowest_chi)` (main.cpp:809)
        best, low = grid[0], float("inf")
        for lam, c in zip(grid, res["chisq"]):
            if c < low:
                low, best = c, lam
        smooth = best
        table = (grid, res["chisq"])
    x, f, fp, ok = fit(flat, smooth, device=device, log_pen=log_pen, **opt)
    out = {"smoothing": smooth, "objective": f, "converged": ok, "cv": table,
           "dated_tree": output.dated_tree(flat, names, fp, x), "rate_tree": output.rate_tree(flat, names, fp, x, numsites),
           "x": x, "freeparams": fp}
    if outfile:
        output.write_results(outfile, flat, names, fp, x, numsites, cv_table=table, cvoutfile=cvoutfile)
    return out
