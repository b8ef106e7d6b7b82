"""Batched truncated Newton (tpl_optimize_batch, method = 1) on the GPU through the C ABI:
  * the CUDA Hessian-vector kernel equals the CPU emulation of the same node function (tests/tn_sim.cpp), which
    tests/test_tn_sim.py pins against finite differences of the oracle's gradient;
  * the optimiser reaches the reference's END-TO-END results and the CPU emulation's optimum;
  * graph replay == plain launches, bitwise; per-vector CV masks and smoothing values are honoured."""
import numpy as np
import pytest

from golden_util import Golden
from treepl_b200 import batchopt, cv
from treepl_b200 import engine as eng
from treepl_b200 import treeflat as tf
import test_tn_sim as TS

pytestmark = pytest.mark.gpu


def _engine(pr, flat):
    plp = eng.PLCalcParallel.from_flat(flat, device=0)
    plp.set_freeparams(pr.P, False, pr.fp)
    masks = None
    if pr.mask is not None:
        masks = [list(np.flatnonzero(pr.mask[b])) for b in range(pr.B)]
    plp.batch_configure(pr.B, lane_smoothing=pr.lams, masks=masks)
    return plp


@pytest.mark.parametrize("name,lams,masked", [("test", [0.1, 10.0], False), ("M1155", [10.0, 0.1, 1000.0, 1.0], True),
                                              ("big_test", [1.0] * 33, True), ("M1155", [10.0, 0.1] * 20, True),
                                              ("pl4203", [100.0] * 17, False)])
def test_hessvec_kernel_equals_cpu_emulation(name, lams, masked):
    """B <= 16: lane-packed tn_hv_kernel; B > 16: the tiled shared-memory kernel (tn_hv_tile_kernel)."""
    sim = TS.load_sim()
    flat = Golden(name).flat
    n = len(flat["parents"])
    tips = np.flatnonzero(flat["child_counts"] == 0)
    masks = None
    if masked:
        masks = [[int(tips[(3 + 2 * b) % len(tips)])] for b in range(len(lams))]
        masks[0] = []
    pr = TS.Problem(flat, lams, masks)
    P, B = pr.P, pr.B
    rng = np.random.RandomState(11)
    x0 = cv.start_vector(flat, P, n - 1)
    X = np.repeat(x0[:, None], B, axis=1)
    X[: n - 1] *= np.exp(0.2 * rng.randn(n - 1, B))
    V = rng.randn(P, B)
    plp = _engine(pr, flat)
    F, G = plp.calc_pl_grad_batch(X, mode=eng.GRAD_AD)
    W, DG = plp.debug_hessvec(X, V, pr.sc)
    Ws, DGs = np.zeros((P, B)), np.zeros((P, B))
    Xc, Gc, Vc = np.ascontiguousarray(X), np.ascontiguousarray(G), np.ascontiguousarray(V)
    sim.tn_sim_hessvec(P, B, *pr.tree_args(), TS._d(pr.sc), TS._d(pr.lams), 0.0025, TS._d(Xc), TS._d(Gc), TS._d(Vc), TS._d(Ws), TS._d(DGs), 0)
    scale = np.abs(Ws).max(axis=0)
    assert np.all(np.abs(W - Ws).max(axis=0) <= 1e-11 * scale), np.abs(W - Ws).max(axis=0) / scale
    assert np.allclose(DG, DGs, rtol=1e-11, atol=0)
    plp.close()


def _run(name, lams, masks=None, **opt):
    flat = Golden(name).flat
    pr = TS.Problem(flat, lams, masks)
    plp = _engine(pr, flat)
    lower, upper = batchopt.reference_bounds(flat, pr.fp, pr.P)
    x0 = cv.start_vector(flat, pr.P, pr.n - 1)
    X0 = np.repeat(x0[:, None], pr.B, axis=1)
    X, F, info = plp.optimize_batch(X0, lower, upper, method=1, date_scale=float(pr.sc[-1]), **opt)
    return pr, plp, X, F, info


@pytest.mark.parametrize("name,lams,masked,log_pen", [("test", [0.1, 100.0], False, False), ("M1155", [10.0, 0.1, 1000.0, 1.0], True, False),
                                                      ("vib", [1e4, 1.0], True, True), ("big_test", [1e-4] * 17 + [100.0] * 17, True, False)])
def test_tree_preconditioner_kernels_equal_cpu_emulation(name, lams, masked, log_pen):
    """zz = M^-1 r through the sweep kernels (tn_tree_up / down per wide level, heavy paths in tn_tree_top_kernel) == the same
    node functions run in elimination order on the CPU (tests/tn_sim.cpp::tree_solve), which tests/test_tn_sim.py pins against a dense solve."""
    sim = TS.load_sim()
    flat = Golden(name).flat
    n = len(flat["parents"])
    tips = np.flatnonzero(flat["child_counts"] == 0)
    masks = None
    if masked:
        masks = [[int(tips[(3 + 2 * b) % len(tips)])] for b in range(len(lams))]
        masks[0] = []
    pr = TS.Problem(flat, lams, masks, log_pen=log_pen)
    P, B = pr.P, pr.B
    rng = np.random.RandomState(12)
    x0 = cv.start_vector(flat, P, n - 1)
    X = np.repeat(x0[:, None], B, axis=1)
    X[: n - 1] *= np.exp(0.1 * rng.randn(n - 1, B))
    rhs = rng.randn(P, B)
    free = (rng.rand(P, B) > 0.02).astype(np.float64)            # a few frozen rows per vector
    plp = _engine(pr, flat)
    plp.set_log_pen(log_pen)
    F, G = plp.calc_pl_grad_batch(X, mode=eng.GRAD_AD)
    rhs2 = rng.randn(P, B)
    # out: factorise + solve (the sweep after an accepted step); out2: the solve-only sweeps of a CG iteration, same factors
    out, out2 = plp.debug_treesolve(X, free, rhs, pr.sc, rhs2=rhs2)
    Xc, Gc, Fc = (np.ascontiguousarray(v) for v in (X, G, free))
    # elimination amplifies the rounding differences between host and device arithmetic (FMA contraction) by the condition of
    # the pivots: 1e-11 typical; up to 1e-5 on the 9,313-node tree at this random, non-optimal point
    tol = 1e-4 if name == "big_test" else 1e-7
    for got, r in ((out, rhs), (out2, rhs2)):
        want = np.zeros((P, B))
        Rc = np.ascontiguousarray(r)
        sim.tn_sim_treesolve(P, B, *pr.tree_args(), TS._d(pr.sc), TS._d(pr.lams), 0.0025, TS._d(Xc), TS._d(Gc), TS._d(Fc), TS._d(Rc), TS._d(want),
                             int(log_pen))
        scale = np.abs(want).max(axis=0)
        assert np.all(np.isfinite(got))
        err = np.abs(got - want).max(axis=0) / scale
        assert np.all(err <= tol) and np.median(err) <= 1e-7, err
        assert np.all(got[free == 0] == 0.0)
    # linearity in the right-hand side with fixed factors: M^-1 (rhs + rhs2) through the solve-only sweeps
    _, out12 = plp.debug_treesolve(X, free, rhs, pr.sc, rhs2=rhs + rhs2)
    s12 = np.abs(out + out2).max(axis=0)
    assert np.all(np.abs(out12 - (out + out2)).max(axis=0) <= 1e-8 * s12)
    plp.close()


@pytest.mark.parametrize("precond", [1, 2])
def test_tn_reproduces_reference_final_pl_and_cpu_emulation(precond):
    """precond 1: Jacobi (round 1); 2: block-tree factorisation (default).  Same optima, the reference's final PL values."""
    sim = TS.load_sim()
    for name, lams, ref in (("test", [0.1, 0.1], 234.23049), ("pl4203", [100.0, 1.0], 390.04656), ("M1155", [10.0, 1000.0, 0.1, 1.0], 252.3558)):
        pr, plp, X, F, info = _run(name, lams, max_steps=20000, precond=precond)
        assert ((info["status"] >= 1) & (info["status"] <= 2)).all(), info
        assert abs(F[0] - ref) <= (1e-4 if name == "test" else 2e-6) * ref and F[0] <= ref + 1e-6 * ref, (name, F)
        _, Xs, Fs, ss, its, ncg, steps = TS.fit(sim, name, lams, precond=precond - 1, cg_per_step=4 if precond == 1 else 2)
        assert np.all(np.abs(F - Fs) <= 1e-9 * np.abs(Fs)), (name, F, Fs)
        assert 0.5 * steps <= info["steps"] <= 2.0 * steps + 64, (info["steps"], steps)
        if precond == 2:
            assert info["cg_iterations"] <= 4 * int(info["iterations"].sum()) + 8, info      # a handful of CG iterations per Newton step
        # a stationary point of the engine's own objective
        f, g = plp.calc_pl_grad_batch(X, mode=eng.GRAD_AD)
        assert np.all(np.abs(f - F) <= 1e-12 * np.abs(F))
        plp.close()


def test_tn_log_penalty_objective_reaches_the_lbfgs_optimum():
    """examples/big_test.cppr8s sets log_pen (pl_calc_parallel.cpp:866-879: the root-children variance term on ln r);
    truncated Newton carries that Hessian block: same optimum as the L-BFGS path and as the CPU emulation."""
    sim = TS.load_sim()
    for name, lams in (("vib", [1000.0, 10.0]), ("test", [0.1, 100.0])):
        flat = Golden(name).flat
        pr = TS.Problem(flat, lams, log_pen=True)
        plp = _engine(pr, flat)
        plp.set_log_pen(True)
        lower, upper = batchopt.reference_bounds(flat, pr.fp, pr.P)
        x0 = cv.start_vector(flat, pr.P, pr.n - 1)
        X0 = np.repeat(x0[:, None], pr.B, axis=1)
        X, F, info = plp.optimize_batch(X0, lower, upper, method=1, date_scale=float(pr.sc[-1]), max_steps=20000)
        Xl, Fl, il = plp.optimize_batch(X0, lower, upper, method=0, date_scale=float(pr.sc[-1]), max_steps=40000, ftol=1e-15, patience=40)
        plp.close()
        assert ((info["status"] >= 1) & (info["status"] <= 2)).all(), info
        assert np.all(np.abs(F - Fl) <= 1e-7 * np.abs(Fl)) and np.all(F <= Fl + 1e-9 * np.abs(Fl)), (name, F, Fl)
        _, Xs, Fs, ss, its, ncg, steps = TS.fit(sim, name, lams, precond=1, cg_per_step=2, log_pen=True)
        assert np.all(np.abs(F - Fs) <= 1e-9 * np.abs(Fs)), (name, F, Fs)


def test_tn_converges_for_every_smoothing_value_on_a_large_tree():
    """the regime round 1 left open: large tree x large smoothing.  20,000-tip synthetic tree, smoothing
    1e4 ... 1e-2: every vector gradient-converged, a few CG iterations per Newton step."""
    ft, _, _ = tf.synth_problem(20000, seed=3, ncal=8)
    flat = ft.as_dict()
    lams = [1e4, 1e3, 300.0, 100.0, 10.0, 1.0, 0.1, 0.01]
    fp, P = tf.generate_param_order_vector(flat["free"], lf=False)
    n = len(flat["parents"])
    plp = eng.PLCalcParallel.from_flat(flat, device=0)
    plp.set_freeparams(P, False, fp)
    plp.batch_configure(len(lams), lane_smoothing=np.array(lams))
    lower, upper = batchopt.reference_bounds(flat, fp, P)
    x0 = cv.start_vector(flat, P, n - 1)
    par = flat["parents"]
    ds = float(np.median(np.maximum(flat["start_dates"][par[1:]] - flat["start_dates"][1:], 1e-12)))
    X, F, info = plp.optimize_batch(np.repeat(x0[:, None], len(lams), axis=1), lower, upper, method=1, date_scale=ds, max_steps=6000, use_graph=1)
    assert ((info["status"] >= 1) & (info["status"] <= 2)).all(), info
    assert np.all(np.diff(F) < 0)                     # less smoothing, lower objective
    f, g = plp.calc_pl_grad_batch(X, mode=eng.GRAD_AD)
    gz = g.copy()
    gz[: n - 1] *= X[: n - 1]
    gz[n - 1:] *= ds
    assert np.all(np.abs(gz).max(axis=0) <= 1e-6 * np.abs(f)), np.abs(gz).max(axis=0) / np.abs(f)
    plp.close()


def test_tn_graph_replay_is_bitwise_and_uses_fewer_evaluations_than_lbfgs():
    lams = [1000.0, 100.0, 10.0, 1.0, 0.1, 0.1]
    runs = []
    for graph in (0, 1, 1):
        pr, plp, X, F, info = _run("M1155", lams, max_steps=20000, use_graph=graph)
        runs.append((X, F, info))
        plp.close()
    assert np.array_equal(runs[0][1], runs[1][1]) and np.array_equal(runs[0][0], runs[1][0])
    assert np.array_equal(runs[1][1], runs[2][1]) and np.array_equal(runs[1][0], runs[2][0])
    flat = Golden("M1155").flat
    pr = TS.Problem(flat, lams)
    plp = _engine(pr, flat)
    lower, upper = batchopt.reference_bounds(flat, pr.fp, pr.P)
    x0 = cv.start_vector(flat, pr.P, pr.n - 1)
    Xl, Fl, il = plp.optimize_batch(np.repeat(x0[:, None], pr.B, axis=1), lower, upper, method=0, date_scale=float(pr.sc[-1]), max_steps=20000,
                                    ftol=1e-15, patience=40)
    plp.close()
    assert np.all(runs[0][1] <= Fl + 1e-9 * np.abs(Fl)), (runs[0][1], Fl)
    assert runs[0][2]["steps"] < 0.7 * il["steps"], (runs[0][2]["steps"], il["steps"])


def test_tn_loocv_m1155_chisq_table():
    gd = Golden("M1155")
    cals = [(t, lo, hi) for t, lo, hi in gd.meta["calibrations"]]
    ft = tf.flatten_newick(gd.meta["newick"], gd.meta["numsites"], cals, seed=1)
    res = cv.loocv(dict(gd.flat), [1000.0, 100.0, 10.0, 1.0, 0.1], ft.tips_postorder, method=1)
    ref = np.array([2730.97, 2365.8, 2130.28, 2178.18, 2187.12])
    assert res["cv_converged"] == res["n_vectors"]
    assert np.all(np.abs(np.array(res["chisq"]) - ref) <= 2e-5 * ref), res["chisq"]
    assert res["best"] == 10.0


def test_multistart_finds_the_common_optimum():
    """random restarts (treepl_b200/cv.py::multistart): every start of M1155 at smoothing 10 reaches the reference's
    final PL 252.3558, and the exchanged best is the minimum over the starts"""
    flat = dict(Golden("M1155").flat)
    res = cv.multistart(flat, 10.0, 16, jitter=0.3, seed=3)
    assert res["local_converged"] == 16
    assert np.all(np.abs(res["local_objectives"] - 252.3558) <= 2e-6 * 252.3558), res["local_objectives"]
    assert res["best"] == res["local_objectives"].min() and res["winner_rank"] == 0


def test_tn_refuses_what_it_does_not_support_loudly():
    flat = dict(Golden("test").flat)
    pr = TS.Problem(flat, [0.1, 0.1])
    lower, upper = batchopt.reference_bounds(flat, pr.fp, pr.P)
    x0 = cv.start_vector(flat, pr.P, pr.n - 1)
    X0 = np.repeat(x0[:, None], 2, axis=1)
    plp = _engine(pr, flat)
    with pytest.raises(eng.EngineError, match="TPL_GRAD_AD"):
        plp.optimize_batch(X0, lower, upper, method=1, mode=eng.GRAD_ANALYTIC)
    with pytest.raises(eng.EngineError, match="method must be"):
        plp.optimize_batch(X0, lower, upper, method=7)
    cvm = np.zeros(pr.n, np.int32)
    cvm[12] = 1
    plp.set_cv_nodes(cvm)
    with pytest.raises(eng.EngineError, match="per vector"):
        plp.optimize_batch(X0, lower, upper, method=1)
    plp.close()
    # Langley-Fitch layout: one shared rate
    plp = eng.PLCalcParallel.from_flat(flat, device=0)
    fp, P = eng.generate_param_order_vector(flat["free"], lf=True)
    x = plp.set_freeparams(P, True, fp)
    with pytest.raises(eng.EngineError, match="PL layout"):
        plp.optimize_batch(np.repeat(x[:, None], 2, axis=1), np.full(P, 1e-40), np.full(P, 1e15), method=1)
    plp.close()


def test_end_to_end_dating_of_the_example_configs(tmp_path):
    """treepl_b200/run.py::date_tree on examples/test.cppr8s (smooth = 0.1) and examples/M1155.cppr8s (cv): the reference's
    dated tree (node heights within 2e-3), selected smoothing, final PL, cv.out lines"""
    import test_output as TO
    from treepl_b200 import output, run
    gd = Golden("test")
    r = run.date_tree(dict(gd.flat), gd.meta["names"], gd.meta["numsites"], smooth=0.1, outfile=str(tmp_path / "out_dates.tre"))
    assert r["converged"] and abs(r["objective"] - 234.23049) <= 1e-4 * 234.23049
    strip = lambda s: "".join(ch for ch in s if not (ch.isdigit() or ch == "."))
    assert strip(r["dated_tree"]) == strip(TO.REF_DATED)
    g = [float(t.split(")")[0].split(",")[0]) for t in r["dated_tree"].split(":")[1:]]
    ref = [float(t.split(")")[0].split(",")[0]) for t in TO.REF_DATED.split(":")[1:]]
    assert np.allclose(g, ref, rtol=0, atol=1.5e-2), np.abs(np.array(g) - np.array(ref)).max()     # the reference stops on the bound
    assert open(tmp_path / "out_dates.tre").read().strip() == r["dated_tree"]
    assert open(str(tmp_path / "out_dates.tre") + ".r8s").read().strip() == r["rate_tree"]
    gd = Golden("M1155")
    r = run.date_tree(dict(gd.flat), gd.meta["names"], gd.meta["numsites"], do_cv=True, outfile=str(tmp_path / "m.tre"),
                      cvoutfile=str(tmp_path / "cv.out"))
    assert r["smoothing"] == 10.0 and abs(r["objective"] - 252.3558) <= 2e-6 * 252.3558
    assert open(tmp_path / "cv.out").read().split("\n")[:5] == ["chisq: (1000) 2730.97", "chisq: (100) 2365.8", "chisq: (10) 2130.28",
                                                              "chisq: (1) 2178.17", "chisq: (0.1) 2187.12"]


def test_random_subsample_cv_end_to_end():
    """examples/pl4203.cppr8s runs a random-subsample CV (cvstart 1e-4 ... cvstop 100, x 0.1 is its own grid; here the default
    grid direction 1000 -> 0.1): folds are sister-free tip groups, every fold optimisation converges, and the final fit at the
    selected smoothing is the optimum of the engine's own objective."""
    from treepl_b200 import run
    gd = Golden("pl4203")
    flat = dict(gd.flat)
    r = run.date_tree(flat, gd.meta["names"], gd.meta["numsites"], do_cv=True, randomcv=True, randomcviter=6, randomcvsamp=0.15, seed=2)
    grid, chisq = r["cv"]
    assert grid == [1000.0, 100.0, 10.0, 1.0, 0.1] and len(chisq) == 5 and np.all(np.isfinite(chisq)) and np.all(np.array(chisq) > 0)
    assert r["smoothing"] == grid[int(np.argmin(chisq))] and r["converged"]
    pr = TS.Problem(flat, [r["smoothing"], r["smoothing"]])
    plp = _engine(pr, flat)
    f, g = plp.calc_pl_grad_batch(np.repeat(r["x"][:, None], 2, axis=1), mode=eng.GRAD_AD)
    assert abs(f[0] - r["objective"]) <= 1e-12 * abs(f[0])
    gz = g[:, 0].copy()
    gz[: pr.nr] *= r["x"][: pr.nr]
    assert np.abs(gz).max() <= 1e-6 * abs(f[0])
    plp.close()


def test_engines_on_two_devices_of_one_process():
    """The C ABI lets one process hold engines on several GPUs (tpl_create's device argument).  Kernel attributes (the pipelined
    Hessian-vector kernel's shared-memory opt-in) and the cooperative-launch state are per device: the same 40-vector
    optimisation on device 0 and device 1 gives bitwise the same optimum.  Needs two visible GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    flat = dict(Golden("pl4203").flat)
    n = len(flat["parents"])
    fp, P = eng.generate_param_order_vector(flat["free"], lf=False)
    lo, hi = batchopt.reference_bounds(flat, fp, P)
    x0 = cv.start_vector(flat, P, n - 1)
    B = 40
    lam = 10.0 ** np.linspace(2, -1, B)
    out = []
    for dev in (1, 0, 1):
        plp = eng.PLCalcParallel.from_flat(flat, device=dev)
        plp.set_freeparams(P, False, fp)
        plp.batch_configure(B, lane_smoothing=lam)
        X, F, info = plp.optimize_batch(np.repeat(x0[:, None], B, axis=1), lo, hi, method=1, max_steps=4000)
        assert ((info["status"] >= 1) & (info["status"] <= 3)).all()
        out.append((X, F))
        plp.close()
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[2][1])
