"""End-to-end check of the batched device-resident optimiser + LOOCV scoring (first cut):
the optimum it reaches (objective, dated nodes, chi-square per fold) against an independent CPU optimisation
of the ORACLE objective (scipy L-BFGS-B) on the reference's example trees."""
import numpy as np
import pytest

from golden_util import Golden
from oracle import plo
from treepl_b200 import batchopt, cv
from treepl_b200 import engine as eng
from treepl_b200 import treeflat as tf

pytestmark = pytest.mark.gpu


def _oracle_fit(flat, lam, tip, x0, lo, hi, nr):
    from scipy.optimize import minimize

    o = plo.OracleProblem(flat)
    n = len(flat["parents"])
    cvm = np.zeros(n, np.int32)
    keep = np.ones(x0.shape[0], bool)
    if tip is not None:
        cvm[tip] = 1
        keep[tip - 1] = False
    o.set_smoothing(lam)
    if tip is not None:
        # the fold engine starts from, and a pruned tip keeps, the full-data optimum (cvstrates, main.cpp:681,702-704);
        # x0 is that optimum in the full layout
        fp_full, _ = tf.generate_param_order_vector(flat["free"], lf=False)
        rates = np.array(flat["start_rates"], dtype=np.float64)
        rates[1:] = x0[fp_full[1:n]]
        o.set_state(rates, flat["start_dates"])
    o.set_cv(cvm if tip is not None else None)
    o.set_freeparams(False, use_cv=tip is not None)
    nrc = nr - (0 if tip is None else 1)

    def fun(z):
        x = z.copy()
        x[:nrc] = np.exp(z[:nrc])
        f, g = o.calc_function_gradient(x)
        if not np.isfinite(f) or f >= 1e15:
            return 1e15, np.zeros_like(z)
        g = g.copy()
        g[:nrc] *= x[:nrc]
        return f, g

    z0 = x0[keep].copy()
    z0[:nrc] = np.log(z0[:nrc])
    lo2, hi2 = lo[keep].copy(), hi[keep].copy()
    lo2[:nrc] = np.log(lo2[:nrc]); hi2[:nrc] = np.log(hi2[:nrc])
    # open date bounds, as the batched optimiser keeps them (exactly ON a penalised bound the objective drops
    # the barrier term: a spurious discontinuous minimum)
    m = 1e-9 * np.maximum(1.0, np.abs(np.where(hi2[nrc:] < 1e19, hi2[nrc:], lo2[nrc:])))
    lo2[nrc:] += m
    hi2[nrc:] -= np.where(hi2[nrc:] < 1e19, m, 0.0)
    best = None
    for _ in range(6):           # restart from the last point until it stops improving
        r = minimize(fun, z0, jac=True, method="L-BFGS-B", bounds=list(zip(lo2, hi2)),
                     options={"maxiter": 20000, "maxfun": 200000, "ftol": 1e-15, "gtol": 1e-10, "maxcor": 30})
        z0 = r.x
        if best is not None and abs(best - r.fun) <= 1e-12 * abs(r.fun):
            break
        best = r.fun
    x = r.x.copy()
    x[:nrc] = np.exp(x[:nrc])
    full = np.zeros(x0.shape[0])
    full[keep] = x
    return r.fun, full


@pytest.mark.parametrize("name,lams", [("test", [0.1, 10.0]), ("pl4203", [1.0, 100.0])])
def test_batched_loocv_reaches_the_oracle_optimum(name, lams):
    gd = Golden(name)
    flat = gd.flat
    cals = [(t, lo, hi) for t, lo, hi in gd.meta["calibrations"]]
    ft = tf.flatten_newick(gd.meta["newick"], gd.meta["numsites"], cals, seed=1)
    tips = ft.tips_postorder
    res = cv.loocv(flat, lams, tips, max_iter=6000)
    n = len(flat["parents"])
    fp, P = tf.generate_param_order_vector(flat["free"], lf=False)
    lo, hi = batchopt.reference_bounds(flat, fp, P)
    x0 = cv.start_vector(flat, P, n - 1)
    for k, lam in enumerate(lams):
        f_full, x_full = _oracle_fit(flat, lam, None, x0, lo, hi, n - 1)
        assert abs(res["full_objective"][k] - f_full) <= 1e-7 * abs(f_full), (lam, res["full_objective"][k], f_full)
        chisq = 0.0
        for tip in tips:
            f, x = _oracle_fit(flat, lam, int(tip), x_full, lo, hi, n - 1)
            chisq += cv.loocv_chisq(flat, x[:, None], [int(tip)], fp)[0]
        assert abs(res["chisq"][k] - chisq) <= 2e-3 * abs(chisq), (lam, res["chisq"][k], chisq)


def test_pruned_tip_on_the_root_keeps_the_full_data_rate():
    """examples/vib: node 174 is a tip that hangs off the root.  Pruned, its rate still enters the root-children
    variance term (pl_calc_parallel.cpp:866-878) and must be the full-data optimum of that smoothing value, as in the
    reference's fold engines (main.cpp:681,702-704) — not the engine's start rate."""
    gd = Golden("vib")
    flat = gd.flat
    cals = [(t, lo, hi) for t, lo, hi in gd.meta["calibrations"]]
    ft = tf.flatten_newick(gd.meta["newick"], gd.meta["numsites"], cals, seed=1)
    n = len(flat["parents"])
    root_tips = [int(t) for t in ft.tips_postorder if flat["parents"][t] == 0]
    assert root_tips, "vib has a tip child of the root"
    tips = root_tips + [int(t) for t in ft.tips_postorder if flat["parents"][t] != 0][:3]
    lams = [1000.0, 10.0]
    res = cv.loocv(flat, lams, tips, max_iter=6000)
    fp, P = tf.generate_param_order_vector(flat["free"], lf=False)
    lo, hi = batchopt.reference_bounds(flat, fp, P)
    x0 = cv.start_vector(flat, P, n - 1)
    for k, lam in enumerate(lams):
        f_full, x_full = _oracle_fit(flat, lam, None, x0, lo, hi, n - 1)
        assert abs(res["full_objective"][k] - f_full) <= 1e-7 * abs(f_full)
        for g, tip in enumerate(tips):
            f, x = _oracle_fit(flat, lam, tip, x_full, lo, hi, n - 1)
            want = cv.loocv_chisq(flat, x[:, None], [tip], fp)[0]
            got = res["terms"][k, g]
            assert abs(got - want) <= 2e-3 * max(abs(want), 1e-3), (lam, tip, got, want)


def test_start_columns_and_probes_equal_the_expanded_batch():
    """tpl_optimize_batch_cols (start vectors = columns of a [P, K] matrix expanded on the device, only probed rows returned)
    == tpl_optimize_batch on the batch expanded by the host, bit for bit: optima, objectives, status, probes."""
    flat = Golden("M1155").flat
    n = len(flat["parents"])
    fp, P = eng.generate_param_order_vector(flat["free"], lf=False)
    lo, hi = batchopt.reference_bounds(flat, fp, P)
    x0 = cv.start_vector(flat, P, n - 1)
    rng = np.random.RandomState(4)
    K, B = 3, 14
    Xs = np.repeat(x0[:, None], K, axis=1)
    Xs[: n - 1] *= np.exp(0.2 * rng.randn(n - 1, K))
    col = rng.randint(0, K, size=B).astype(np.int32)
    lam = 10.0 ** rng.uniform(-1, 3, size=B)
    tips = np.flatnonzero(flat["child_counts"] == 0)
    masks = [[int(tips[(5 + 3 * b) % len(tips)])] if b % 3 else [] for b in range(B)]
    rows = rng.randint(-1, P, size=(5, B)).astype(np.int32)
    par = flat["parents"]
    dscale = float(np.median(np.maximum(flat["start_dates"][par[1:]] - flat["start_dates"][1:], 1e-12)))
    plp = eng.PLCalcParallel.from_flat(flat, device=0)
    plp.set_freeparams(P, False, fp)
    plp.batch_configure(B, lane_smoothing=lam, masks=masks)
    X1, F1, i1 = plp.optimize_batch(np.ascontiguousarray(Xs[:, col]), lo, hi, method=1, date_scale=dscale, max_steps=4000)
    F2, probes, X2, i2 = plp.optimize_batch_cols(Xs, col, lo, hi, probe_rows=rows, want_x=True, date_scale=dscale, max_steps=4000)
    assert ((i1["status"] >= 1) & (i1["status"] <= 3)).all()
    assert np.array_equal(F1, F2) and np.array_equal(X1, X2) and np.array_equal(i1["status"], i2["status"])
    want = np.where(rows >= 0, X1[np.maximum(rows, 0), np.arange(B)[None, :]], 0.0)
    assert np.array_equal(probes, want)
    # without probes and without X only the objectives return
    F3, p3, X3, _ = plp.optimize_batch_cols(Xs, col, lo, hi, date_scale=dscale, max_steps=4000)
    assert np.array_equal(F3, F1) and p3.shape == (0, B) and X3 is None
    with pytest.raises(eng.EngineError):
        plp.optimize_batch_cols(Xs, col + K, lo, hi, date_scale=dscale)            # column index out of range
    with pytest.raises(eng.EngineError):
        plp.optimize_batch_cols(Xs, col, lo, hi, method=0)                         # L-BFGS has no start-column entry
    plp.close()
