"""Device-resident batched L-BFGS (tpl_optimize_batch) on the GPU, through the C ABI:
  * the optimum it reaches is the reference's own END-TO-END result (final PL, dated tree, LOOCV chi-square table of
    the `seed=1, nthreads=1` reference runs);
  * it follows the CPU emulation of the same state machine (tests/lbfgs_sim.cpp + oracle) to the optimum;
  * CUDA-graph replay and plain launches give bitwise identical results; repeated runs are bitwise identical."""
import numpy as np
import pytest

from golden_util import Golden
from treepl_b200 import batchopt, cv
from treepl_b200 import engine as eng
from treepl_b200 import treeflat as tf

pytestmark = pytest.mark.gpu


def _setup(name, lams, masks=None):
    gd = Golden(name)
    flat = gd.flat
    n = len(flat["parents"])
    plp = eng.PLCalcParallel.from_flat(flat, device=0)
    fp, P = eng.generate_param_order_vector(flat["free"], lf=False)
    plp.set_freeparams(P, False, fp)
    lo, hi = batchopt.reference_bounds(flat, fp, P)
    B = len(lams)
    plp.batch_configure(B, lane_smoothing=np.array(lams, dtype=np.float64), masks=masks)
    x0 = cv.start_vector(flat, P, n - 1)
    par = flat["parents"]
    ds = float(np.median(np.maximum(flat["start_dates"][par[1:]] - flat["start_dates"][1:], 1e-12)))
    return flat, plp, fp, P, lo, hi, np.repeat(x0[:, None], B, axis=1), ds


def test_native_optimizer_reproduces_reference_final_pl_and_dated_tree():
    flat, plp, fp, P, lo, hi, X0, ds = _setup("test", [0.1, 0.1])
    X, F, info = plp.optimize_batch(X0, lo, hi, date_scale=ds, max_steps=4000)
    assert ((info["status"] >= 1) & (info["status"] <= 3)).all(), info
    assert abs(F[0] - 234.23049) <= 1e-4 * 234.23049 and F[0] <= 234.23049
    assert F[0] == F[1]
    n = 13
    dates = {i: X[fp[n + i], 0] for i in range(n) if fp[n + i] != -1}
    want = {1: 6.626589, 4: 10 - 2.000579, 5: 4.807152, 7: 3.207058, 10: 3.997337}
    for node, h in want.items():
        assert abs(dates[node] - h) <= 2e-3 * h, (node, dates[node], h)
    # the optimum is a stationary point of the engine's own objective
    f, g = plp.calc_pl_grad_batch(X, mode=eng.GRAD_AD)
    assert abs(f[0] - F[0]) <= 1e-12 * abs(F[0])
    plp.close()


@pytest.mark.parametrize("history", [4, 8, 16])
def test_native_optimizer_final_pl_of_example_configs(history):
    """pl4203 at smoothing 100: 390.04656; M1155 at smoothing 10: 252.3558 (reference final PL)"""
    flat, plp, fp, P, lo, hi, X0, ds = _setup("pl4203", [100.0, 1.0])
    X, F, info = plp.optimize_batch(X0, lo, hi, date_scale=ds, history=history, max_steps=8000)
    assert ((info["status"] >= 1) & (info["status"] <= 3)).all(), info
    assert abs(F[0] - 390.04656) <= 2e-6 * 390.04656, F
    plp.close()
    flat, plp, fp, P, lo, hi, X0, ds = _setup("M1155", [10.0, 1000.0])
    X, F, info = plp.optimize_batch(X0, lo, hi, date_scale=ds, history=history, max_steps=10000)
    assert ((info["status"] >= 1) & (info["status"] <= 3)).all(), info
    assert abs(F[0] - 252.3558) <= 2e-6 * 252.3558, F
    plp.close()


def test_native_optimizer_follows_the_cpu_emulation():
    """same state machine (tpl_lbfgs_core.h), oracle objective on the CPU vs PL kernels on the GPU: same optimum,
    similar effort"""
    import test_lbfgs_sim as T

    lib = T.load_sim()
    lams = [1000.0, 10.0, 0.1, 1.0]
    _, _, Xs, Fs, ss, its, steps_s = T.fit(lib, "M1155", lams, 20000, 8)
    flat, plp, fp, P, lo, hi, X0, ds = _setup("M1155", lams)
    X, F, info = plp.optimize_batch(X0, lo, hi, date_scale=ds, history=8, max_steps=20000)
    assert ((info["status"] >= 1) & (info["status"] <= 3)).all(), info
    assert np.all(np.abs(F - Fs) <= 1e-8 * np.abs(Fs)), (F, Fs)
    assert 0.5 * steps_s <= info["steps"] <= 2.0 * steps_s + 64, (info["steps"], steps_s)
    plp.close()


def test_graph_replay_equals_plain_launches_and_runs_are_deterministic():
    flat, plp, fp, P, lo, hi, X0, ds = _setup("big_test", [1.0] * 32 + [100.0] * 32)
    rng = np.random.RandomState(3)
    n = len(flat["parents"])
    X0 = X0.copy()
    X0[: n - 1] *= np.exp(0.05 * rng.randn(n - 1, 64))            # 64 different starts
    runs = []
    for graph in (0, 1, 1):
        X, F, info = plp.optimize_batch(X0, lo, hi, date_scale=ds, history=8, max_steps=256, check_every=32, use_graph=graph)
        runs.append((X, F, info))
    assert info["steps"] == 256            # clamped to max_steps (the last chunk runs as plain launches)
    assert np.array_equal(runs[0][1], runs[1][1]) and np.array_equal(runs[0][0], runs[1][0])
    assert np.array_equal(runs[1][1], runs[2][1]) and np.array_equal(runs[1][0], runs[2][0])
    # every vector went downhill and stayed feasible
    F0, _ = plp.calc_pl_grad_batch(X0, mode=eng.GRAD_AD)
    assert np.all(runs[0][1] < F0) and np.all(np.isfinite(runs[0][1]))
    Fchk, _ = plp.calc_pl_grad_batch(runs[0][0], mode=eng.GRAD_AD)
    assert np.all(np.abs(Fchk - runs[0][1]) <= 1e-12 * np.abs(Fchk))
    plp.close()


def test_infeasible_start_is_reported_per_vector():
    flat, plp, fp, P, lo, hi, X0, ds = _setup("test", [0.1, 0.1])
    n = 13
    X0 = X0.copy()
    X0[fp[n + 5], 1] = 9.0            # node 5 older than its parent (node 4, 6.77): negative duration -> 1e15
    X, F, info = plp.optimize_batch(X0, lo, hi, date_scale=ds, max_steps=2000)
    assert info["status"][1] == 4 and 1 <= info["status"][0] <= 3, info
    assert abs(F[0] - 234.22339) <= 1e-5 * 234.22339
    plp.close()


def test_loocv_m1155_chisq_table_matches_the_reference_run():
    """examples/M1155.cppr8s: chi-square per smoothing value of the reference run (seed 1, nthreads 1):
    2730.97, 2365.8, 2130.28, 2178.18, 2187.12 -> selected smoothing 10"""
    gd = Golden("M1155")
    cals = [(t, lo, hi) for t, lo, hi in gd.meta["calibrations"]]
    ft = tf.flatten_newick(gd.meta["newick"], gd.meta["numsites"], cals, seed=1)
    res = cv.loocv(dict(gd.flat), [1000.0, 100.0, 10.0, 1.0, 0.1], ft.tips_postorder, max_iter=6000, method=0)
    ref = np.array([2730.97, 2365.8, 2130.28, 2178.18, 2187.12])
    assert res["cv_converged"] == res["n_vectors"]
    assert np.all(np.abs(np.array(res["chisq"]) - ref) <= 5e-5 * ref), res["chisq"]      # the reference prints 6 digits; at smoothing 0.1 chi-square moves by ~2e-5 between equally converged optima
    assert res["best"] == 10.0
