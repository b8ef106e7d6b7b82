"""GPU parity tests: the sm_100a engine, called through the C ABI, against the reference goldens
and the CPU oracle.  Gates (BASELINE.json north_star): objective 1e-10 relative, gradient 1e-8
relative (norm-wise floor of SURVEY.md §7: |dg_i| / max(|g_i|, 1e-8 ||g||_inf)).  The kernels
are far inside those gates; the asserted bounds below are the tighter ones actually achieved."""
import math

import numpy as np
import pytest

from golden_util import Golden, golden_names, rel_err
from oracle import plo
from treepl_b200 import engine as eng
from treepl_b200 import treeflat as tf

pytestmark = pytest.mark.gpu

F_TOL = 1e-12      # gate: 1e-10
G_TOL = 1e-10      # gate: 1e-8


def close_f(got, want, tol=F_TOL):
    if not math.isfinite(want):
        return got == want or (math.isnan(got) and math.isnan(want))
    return abs(got - want) <= tol * max(1.0, abs(want))


def test_fast_math_accuracy():
    rng = np.random.RandomState(1)
    x = np.concatenate([np.exp(rng.uniform(-200, 200, 200000)), rng.uniform(0.5, 2.0, 200000),
                        1.0 + rng.uniform(-1e-3, 1e-3, 100000), [1.0, 2.0, 0.5, 1e-300, 1e300, 5e-324, 0.0, np.inf]])
    lo, rc = eng.debug_math(x)
    with np.errstate(divide="ignore"):
        want_lo, want_rc = np.log(x), 1.0 / x
    fin = np.isfinite(want_lo)
    assert np.array_equal(lo[~fin], want_lo[~fin])
    # absolute error of log: a few 1e-16 (relative to max(1,|log|))
    err = np.abs(lo[fin] - want_lo[fin]) / np.maximum(1.0, np.abs(want_lo[fin]))
    assert err.max() < 4.5e-16, err.max()
    finr = np.isfinite(want_rc) & (want_rc != 0)
    errr = np.abs(rc[finr] - want_rc[finr]) / np.abs(want_rc[finr])
    assert errr.max() < 4.5e-16, errr.max()
    assert np.array_equal(rc[~finr], want_rc[~finr])


@pytest.mark.parametrize("name", golden_names())
def test_single_point_api_matches_reference_goldens(name):
    """calc_pl / calc_gradient / calc_function_gradient on every recorded reference evaluation."""
    gd = Golden(name)
    plp = eng.PLCalcParallel.from_flat(gd.flat)
    assert plp.device_sm() >= 1000
    for e in gd.evals:
        plp.set_log_pen(bool(e["log_pen"]))
        plp.set_state(gd.flat["start_rates"], gd.flat["start_dates"])
        plp.smoothing = e["smoothing"]
        cv = gd.arr(e["k"], "cv")
        plp.set_cv_nodes(cv)
        fp = gd.arr(e["k"], "fp")
        plp.set_freeparams(e["numparams"], bool(e["lf"]), fp)
        x = gd.arr(e["k"], "x")
        want = float(e["f_pl"])
        f = plp.calc_pl(x)
        if e["label"] == "zero_rate" and want > 1e15:
            assert close_f(f, want, 1e-9), (name, e, f)          # sums of LARGE terms
        else:
            assert close_f(f, want), (name, e, f)
        g_an = gd.arr(e["k"], "g_an")
        if g_an is not None:
            g = plp.calc_gradient(x)
            assert rel_err(g, g_an) < G_TOL, (name, e["k"], e["label"], rel_err(g, g_an))
            dur = gd.arr(e["k"], "durations")
            if dur is not None and want < 1e15:
                got = plp.durations
                assert np.array_equal(got[1:], dur[1:])
            # fused form gives the same numbers
            f2, g2 = plp.calc_pl_grad(x, eng.GRAD_ANALYTIC)
            assert f2 == f and np.array_equal(g2, g, equal_nan=True)
        if "f_ad" in e:
            want_ad = float(e["f_ad"])
            f_ad, g_ad = plp.calc_function_gradient(x)
            if want_ad > 1e15 and math.isfinite(want_ad):
                assert close_f(f_ad, want_ad, 1e-9), (name, e, f_ad)
            else:
                assert close_f(f_ad, want_ad), (name, e, f_ad)
            ref_g = gd.arr(e["k"], "g_ad")
            if ref_g is not None and math.isfinite(want_ad) and want_ad < 1e15:
                assert rel_err(g_ad, ref_g) < G_TOL, (name, e["k"], e["label"], rel_err(g_ad, ref_g))
            if want_ad == 1e15:
                assert np.isnan(g_ad).all()       # reference leaves g untouched on the early return
    plp.close()


def _fold_rates(flat, x_full, masked):
    """member rates of the reference's per-fold engine: a pruned node keeps the rate it is handed (main.cpp:702-704);
    in the batched layout that rate travels in the node's own X row."""
    rates = np.array(flat["start_rates"], dtype=np.float64)
    for node in masked:
        rates[node] = x_full[node - 1]
    return rates


def _compact(x_full, n, masked):
    """full-layout vector -> the reference's CV-compacted layout (rate rows of masked nodes removed)."""
    keep = np.ones(x_full.shape[0], dtype=bool)
    for node in masked:
        keep[node - 1] = False
    return x_full[keep], keep


@pytest.mark.parametrize("name,B", [("test", 1), ("test", 5), ("M1155", 32), ("vib", 64), ("clock", 100), ("big_test", 96)])
@pytest.mark.parametrize("mode", [eng.GRAD_ANALYTIC, eng.GRAD_AD])
def test_batched_api_matches_oracle_per_vector(name, B, mode):
    """B vectors with their own smoothing and CV mask in one launch == B oracle evaluations."""
    gd = Golden(name)
    rng = np.random.RandomState(B + 17 * mode)
    n = gd.flat["parents"].shape[0]
    tips = np.flatnonzero(gd.flat["child_counts"] == 0)  # incl. tips that hang off the root (LOOCV prunes them too)
    plp = eng.PLCalcParallel.from_flat(gd.flat)
    o = plo.OracleProblem(gd.flat)
    fp, P = eng.generate_param_order_vector(gd.flat["free"], lf=False)
    x0 = plp.set_freeparams(P, False, fp)
    nr = n - 1
    X = np.empty((P, B))
    base = gd.flat["start_rates"][1] * gd.meta["numsites"]
    for b in range(B):
        X[:nr, b] = base * np.exp(0.4 * rng.randn(nr))
        X[nr:, b] = x0[nr:] * (1 - 1e-4 * rng.rand(P - nr))
    lam = 10.0 ** rng.uniform(-3, 3, B)
    masks = [sorted(rng.choice(tips, rng.randint(0, 1 + len(tips) // 8), replace=False).tolist()) for _ in range(B)]
    if B > 1:
        X[nr + 1, 1] = X[nr + 1, 1] * 1e3 + 1e6      # vector 1 infeasible (bound / negative duration)
    plp.batch_configure(B, lane_smoothing=lam, masks=masks)
    F, G = plp.calc_pl_grad_batch(X, mode=mode)
    F2, G2 = plp.calc_pl_grad_batch(X, mode=mode)
    assert np.array_equal(F, F2) and np.array_equal(G, G2, equal_nan=True)      # deterministic
    Fo, _ = plp.calc_pl_grad_batch(X, mode=mode, want_grad=False)
    assert np.array_equal(F, Fo)
    for b in range(B):
        cv = np.zeros(n, np.int32)
        cv[masks[b]] = 1
        o.set_state(_fold_rates(gd.flat, X[:, b], masks[b]), gd.flat["start_dates"])
        o.set_smoothing(lam[b])
        o.set_cv(cv)
        o.set_freeparams(False, use_cv=True)
        xb, keep = _compact(X[:, b], n, masks[b])
        if mode == eng.GRAD_ANALYTIC:
            f, g = o.calc_pl_and_gradient(xb)
        else:
            f, g = o.calc_function_gradient(xb)
        assert close_f(F[b], f), (name, b, F[b], f)
        if f < 1e15:
            assert rel_err(G[keep, b], g) < G_TOL, (name, b, rel_err(G[keep, b], g))
            assert (G[~keep, b] == 0).all()
    assert B == 1 or F[1] == 1e15
    plp.close()


def test_large_tree_properties_100k_nodes():
    """Size-independent properties at a size the CPU oracle still handles in seconds:
    batch == single point, determinism, lane independence, directional derivative."""
    ft, nw, cals = tf.synth_problem(50000, seed=11, ncal=8)
    flat = ft.as_dict()
    n = ft.numnodes
    plp = eng.PLCalcParallel.from_flat(flat)
    o = plo.OracleProblem(flat)
    fp, P = eng.generate_param_order_vector(ft.free, lf=False)
    x0 = plp.set_freeparams(P, False, fp)
    o.set_freeparams(False)
    plp.smoothing = 100.0
    o.set_smoothing(100.0)
    rng = np.random.RandomState(5)
    B = 64
    X = np.empty((P, B))
    base = ft.start_rates[1] * ft.numsites
    for b in range(B):
        X[: n - 1, b] = base * np.exp(0.3 * rng.randn(n - 1))
        X[n - 1:, b] = x0[n - 1:]
    plp.batch_configure(B)
    F, G = plp.calc_pl_grad_batch(X, mode=eng.GRAD_AD)
    assert plp.last_kernel_generation() == 4          # the bench configuration runs the descriptor ring kernel
    for b in (0, 17, 63):
        f, g = o.calc_function_gradient(X[:, b])
        assert close_f(F[b], f), (b, F[b], f)
        assert rel_err(G[:, b], g) < G_TOL
        f1, g1 = plp.calc_function_gradient(X[:, b].copy())
        # the B=1 launch tiles the node range differently, so the objective's partial sums are
        # grouped differently (a last-bits effect); gradient rows come from the same operands
        assert close_f(f1, F[b], 1e-14) and rel_err(g1, G[:, b]) < G_TOL
    # lane independence: permuting the vectors permutes the results bitwise
    perm = rng.permutation(B)
    Fp, Gp = plp.calc_pl_grad_batch(np.ascontiguousarray(X[:, perm]), mode=eng.GRAD_AD)
    assert np.array_equal(Fp, F[perm]) and np.array_equal(Gp, G[:, perm])
    # directional derivative (central difference) against the gradient; f ~ 5e9 limits the finite
    # difference to ~1e-4 relative, so this only guards against gross errors (a wrong sign, a missing term)
    v = rng.randn(P)
    v /= np.linalg.norm(v)
    x = X[:, 0]
    best = np.inf
    for scale in (1e-5, 1e-4, 1e-3):
        h = scale * np.abs(x).mean()
        fplus = plp.calc_pl_grad(x + h * v, eng.GRAD_AD, want_grad=False)[0]
        fminus = plp.calc_pl_grad(x - h * v, eng.GRAD_AD, want_grad=False)[0]
        dd = (fplus - fminus) / (2 * h)
        best = min(best, abs(dd - G[:, 0] @ v) / max(1.0, abs(dd)))
    assert best <= 1e-3, best
    plp.close()


def test_headline_bench_configuration_against_the_oracle():
    """The EXACT problem and vectors bench.py times (synthetic 100,000-tip tree, N = 199,999, P = 299,996, 16 bounded
    calibrations, B = 256, descriptor ring kernel) against oracle/pl_oracle.c: both gradient flavours, three of the bench's
    own vectors, one CV-masked vector with its own smoothing value and one infeasible vector."""
    import bench
    flat = bench.build_problem(100000)
    n = flat["parents"].shape[0]
    fp, P = eng.generate_param_order_vector(flat["free"], lf=False)
    assert (n, P) == (199999, 299996)
    B = 256
    X = bench.make_vectors(flat, P, B, 1000)
    tips = np.flatnonzero(flat["child_counts"] == 0)
    rng = np.random.RandomState(3)
    masked = np.sort(rng.choice(tips, 10000, replace=False))       # one random-subsample fold of BASELINE config 5
    X[:, 9] = X[:, 8]
    X[n - 1 + 12345, 200] = -1.0                                   # a negative date: infeasible
    lam = np.ones(B)
    lam[9] = 37.0
    plp = eng.PLCalcParallel.from_flat(flat)
    plp.set_freeparams(P, False, fp)
    plp.smoothing = 1.0
    plp.batch_configure(B, lane_smoothing=lam, masks=[masked if b == 9 else [] for b in range(B)])
    o = plo.OracleProblem(flat)
    for mode in (eng.GRAD_AD, eng.GRAD_ANALYTIC):
        F, G = plp.calc_pl_grad_batch(X, mode=mode)
        assert plp.last_kernel_generation() == 4
        assert F[200] == 1e15
        for b in (0, 101, 255, 9):
            cv = np.zeros(n, np.int32)
            keep = np.ones(P, dtype=bool)
            if b == 9:
                cv[masked] = 1
                keep[masked - 1] = False
            o.set_state(_fold_rates(flat, X[:, b], masked if b == 9 else []), flat["start_dates"])
            o.set_smoothing(lam[b])
            o.set_cv(cv)
            o.set_freeparams(False, use_cv=True)
            xb = np.ascontiguousarray(X[keep, b])
            f, g = o.calc_function_gradient(xb) if mode == eng.GRAD_AD else o.calc_pl_and_gradient(xb)
            assert close_f(F[b], f), (mode, b, F[b], f)
            assert rel_err(G[keep, b], g) < G_TOL, (mode, b, rel_err(G[keep, b], g))
            assert (G[~keep, b] == 0).all()
        assert F[9] != F[8]
    plp.close()


def test_state_readback_and_freeparams_emission():
    """evaluate, then read back rates/dates/durations (main.cpp:865-878) and switch LF -> PL (main.cpp:828-833)."""
    gd = Golden("pl4203")
    plp = eng.PLCalcParallel.from_flat(gd.flat)
    o = plo.OracleProblem(gd.flat)
    fp, P = eng.generate_param_order_vector(gd.flat["free"], lf=True)
    x0 = plp.set_freeparams(P, True, fp)
    x0o = o.set_freeparams(True)
    assert np.array_equal(x0, x0o)
    x = x0.copy()
    x[0] *= 1.37
    assert close_f(plp.calc_pl(x), o.calc_pl(x))
    assert np.array_equal(plp.rates, o.doubles("rates"))
    assert np.array_equal(plp.dates, o.doubles("dates"))
    assert np.array_equal(plp.durations[1:], o.doubles("durations")[1:])
    fp2, P2 = eng.generate_param_order_vector(gd.flat["free"], lf=False)
    x1 = plp.set_freeparams(P2, False, fp2)
    x1o = o.set_freeparams(False)
    assert np.array_equal(x1, x1o)           # PL start = LF optimum scattered to every branch
    plp.close()


def test_errors_are_loud():
    gd = Golden("test")
    plp = eng.PLCalcParallel.from_flat(gd.flat)
    with pytest.raises(eng.EngineError):
        plp._ret(plp._L.tpl_calc_pl(plp._h, np.zeros(1).ctypes.data_as(eng.C.POINTER(eng.C.c_double))))
    fp, P = eng.generate_param_order_vector(gd.flat["free"], lf=False)
    plp.set_freeparams(P, False, fp)
    with pytest.raises(ValueError):
        plp.calc_pl(np.zeros(P + 1))
    with pytest.raises(eng.EngineError):
        plp.calc_gradient(np.zeros(P))       # no preceding calc_pl
    plp.close()


@pytest.mark.parametrize("name,B", [("big_test", 64), ("big_test", 34), ("clock", 128), ("M1155", 32)])
@pytest.mark.parametrize("mode", [eng.GRAD_ANALYTIC, eng.GRAD_AD])
@pytest.mark.parametrize("log_pen", [False, True])
def test_tile_kernel_equals_generic_kernel(name, B, mode, log_pen):
    """Generation 2 (shared-memory tiles, cp.async.bulk) against generation 1 on the same vectors, with
    per-vector smoothing and CV masks, and against the oracle for three vectors."""
    gd = Golden(name)
    rng = np.random.RandomState(B + 3 * mode + int(log_pen))
    n = gd.flat["parents"].shape[0]
    tips = np.flatnonzero(gd.flat["child_counts"] == 0)
    plp = eng.PLCalcParallel.from_flat(gd.flat)
    plp.set_log_pen(log_pen)
    fp, P = eng.generate_param_order_vector(gd.flat["free"], lf=False)
    x0 = plp.set_freeparams(P, False, fp)
    nr = n - 1
    base = gd.flat["start_rates"][1] * gd.meta["numsites"]
    X = np.empty((P, B))
    X[:nr, :] = base * np.exp(0.4 * rng.randn(nr, B))
    X[nr:, :] = x0[nr:, None] * (1 - 1e-4 * rng.rand(P - nr, B))
    X[nr + 2, 3] = -1.0                                   # one infeasible vector
    lam = 10.0 ** rng.uniform(-3, 3, B)
    for use_masks in (False, True):
        masks = [sorted(rng.choice(tips, rng.randint(0, 1 + len(tips) // 8), replace=False).tolist())
                 for _ in range(B)] if use_masks else None
        plp.batch_configure(B, lane_smoothing=lam, masks=masks)
        plp.force_kernel(1)
        F1, G1 = plp.calc_pl_grad_batch(X, mode=mode)
        assert plp.last_kernel_generation() == 1
        if not (log_pen and mode == eng.GRAD_ANALYTIC):      # descriptor kernel (no analytic log_pen fast path)
            plp.force_kernel(4)
            F4, G4 = plp.calc_pl_grad_batch(X, mode=mode)
            assert plp.last_kernel_generation() == 4
            F4b, G4b = plp.calc_pl_grad_batch(X, mode=mode)
            assert np.array_equal(F4, F4b) and np.array_equal(G4, G4b, equal_nan=True)     # deterministic
            F4o, _ = plp.calc_pl_grad_batch(X, mode=mode, want_grad=False)
            assert np.array_equal(F4, F4o)
            assert F4[3] == 1e15
            okv = np.arange(B) != 3
            for b in np.flatnonzero(okv):
                assert close_f(F4[b], F1[b], 1e-13), (b, F4[b], F1[b])
        plp.force_kernel(3)                                  # persistent ring kernel
        F3, G3 = plp.calc_pl_grad_batch(X, mode=mode)
        assert plp.last_kernel_generation() == 3
        F3b, G3b = plp.calc_pl_grad_batch(X, mode=mode)
        assert np.array_equal(F3, F3b) and np.array_equal(G3, G3b, equal_nan=True)     # deterministic
        plp.force_kernel(2)
        F2, G2 = plp.calc_pl_grad_batch(X, mode=mode)
        assert plp.last_kernel_generation() == 2
        assert np.array_equal(G3, G2, equal_nan=True)        # same row code: gradient rows bitwise equal
        if not (log_pen and mode == eng.GRAD_ANALYTIC):
            okv = np.arange(B) != 3
            assert np.max(np.abs(F4[okv] - F2[okv]) / np.abs(F2[okv])) < 1e-13
            for b in np.flatnonzero(okv):
                assert rel_err(G4[:, b], G2[:, b]) < G_TOL, (b, rel_err(G4[:, b], G2[:, b]))
        assert F3[3] == 1e15 and np.max(np.abs(F3[np.arange(B) != 3] - F2[np.arange(B) != 3]) / np.abs(F2[np.arange(B) != 3])) < 1e-13
        F2b, G2b = plp.calc_pl_grad_batch(X, mode=mode)
        assert np.array_equal(F2, F2b) and np.array_equal(G2, G2b, equal_nan=True)     # deterministic
        F2o, _ = plp.calc_pl_grad_batch(X, mode=mode, want_grad=False)
        assert np.array_equal(F2, F2o)
        assert F1[3] == 1e15 and F2[3] == 1e15
        ok = np.arange(B) != 3
        assert np.max(np.abs(F1[ok] - F2[ok]) / np.abs(F1[ok])) < 1e-13
        for b in np.flatnonzero(ok):
            assert rel_err(G2[:, b], G1[:, b]) < G_TOL, (b, rel_err(G2[:, b], G1[:, b]))
        o = plo.OracleProblem(gd.flat, log_pen=log_pen)
        for b in (0, B // 2, B - 1):
            cv = np.zeros(n, np.int32)
            if use_masks:
                cv[masks[b]] = 1
            o.set_state(_fold_rates(gd.flat, X[:, b], masks[b] if use_masks else []), gd.flat["start_dates"])
            o.set_smoothing(lam[b])
            o.set_cv(cv)
            o.set_freeparams(False, use_cv=True)
            xb, keep = _compact(X[:, b], n, masks[b] if use_masks else [])
            f, g = o.calc_pl_and_gradient(xb) if mode == eng.GRAD_ANALYTIC else o.calc_function_gradient(xb)
            assert close_f(F2[b], f), (b, F2[b], f)
            assert rel_err(G2[keep, b], g) < G_TOL
    plp.force_kernel(0)
    plp.close()


def test_polytomy_and_unary_nodes_use_the_csr_child_path():
    """Trees that are not strictly binary (collapse, utils.cpp:138-163) must still work in both kernels."""
    nw = "((a:0.1,b:0.2,c:0.15,(d:0.1,e:0.1):0.05):0.1,(f:0.3,(g:0.1,h:0.1,i:0.1):0.2):0.1,j:0.5);"
    ft = tf.flatten_newick(nw, 1000, [(["a", "j"], 50, 50), (["g", "h"], 5, 20)], seed=2)
    flat = ft.as_dict()
    plp = eng.PLCalcParallel.from_flat(flat)
    o = plo.OracleProblem(flat)
    fp, P = eng.generate_param_order_vector(ft.free, lf=False)
    x0 = plp.set_freeparams(P, False, fp)
    o.set_freeparams(False)
    o.set_smoothing(3.0)
    plp.smoothing = 3.0
    rng = np.random.RandomState(4)
    B = 32
    X = np.repeat(x0[:, None], B, axis=1)
    X[: ft.numnodes - 1, :] = 2.0 * np.exp(0.3 * rng.randn(ft.numnodes - 1, B))
    plp.batch_configure(B)
    for gen in (1, 2, 3, 4):
        plp.force_kernel(gen)
        for mode in (eng.GRAD_ANALYTIC, eng.GRAD_AD):
            F, G = plp.calc_pl_grad_batch(X, mode=mode)
            assert plp.last_kernel_generation() == gen
            for b in (0, 31):
                xb = np.ascontiguousarray(X[:, b])
                f, g = o.calc_pl_and_gradient(xb) if mode == eng.GRAD_ANALYTIC else o.calc_function_gradient(xb)
                assert close_f(F[b], f) and rel_err(G[:, b], g) < G_TOL
    plp.close()


@pytest.mark.parametrize("B", [1, 2, 7, 33, 64])
@pytest.mark.parametrize("mode", [eng.GRAD_ANALYTIC, eng.GRAD_AD])
def test_langley_fitch_batches_and_ragged_batch_sizes(B, mode):
    """LF phase (one shared rate, main.cpp:499-552) in batches, odd / tiny batch sizes (generic kernel)."""
    gd = Golden("vib")
    rng = np.random.RandomState(B)
    plp = eng.PLCalcParallel.from_flat(gd.flat)
    o = plo.OracleProblem(gd.flat)
    fp, P = eng.generate_param_order_vector(gd.flat["free"], lf=True)
    x0 = plp.set_freeparams(P, True, fp)
    o.set_freeparams(True)
    X = np.repeat(x0[:, None], B, axis=1)
    X[0, :] = 0.01 * np.exp(rng.randn(B))
    X[1:, :] *= 1 - 1e-3 * rng.rand(P - 1, B)
    plp.batch_configure(B)
    F, G = plp.calc_pl_grad_batch(X, mode=mode)
    assert plp.last_kernel_generation() == 1
    for b in range(B):
        xb = np.ascontiguousarray(X[:, b])
        f, g = o.calc_pl_and_gradient(xb) if mode == eng.GRAD_ANALYTIC else o.calc_function_gradient(xb)
        assert close_f(F[b], f), (b, F[b], f)
        assert rel_err(G[:, b], g) < G_TOL, (b, rel_err(G[:, b], g))
    # PL on the same engine with the same ragged B
    fp, P = eng.generate_param_order_vector(gd.flat["free"], lf=False)
    x0 = plp.set_freeparams(P, False, fp)
    o.set_freeparams(False)
    X = np.repeat(x0[:, None], B, axis=1)
    n = gd.flat["parents"].shape[0]
    X[: n - 1, :] = 0.01 * np.exp(0.3 * rng.randn(n - 1, B))
    plp.batch_configure(B)
    F, G = plp.calc_pl_grad_batch(X, mode=mode)
    for b in range(B):
        xb = np.ascontiguousarray(X[:, b])
        f, g = o.calc_pl_and_gradient(xb) if mode == eng.GRAD_ANALYTIC else o.calc_function_gradient(xb)
        assert close_f(F[b], f) and rel_err(G[:, b], g) < G_TOL
    plp.close()


def test_smallest_trees():
    """Two- and three-tip trees: every node is the root or a child of the root (all rows take the general path)."""
    for nw, cals in (("(a:0.1,b:0.2):0.0;", [(["a", "b"], 10, 10)]),
                     ("((a:0.1,b:0.2):0.05,c:0.3):0.0;", [(["a", "c"], 5, 5)])):
        ft = tf.flatten_newick(nw, 100, cals, seed=1)
        flat = ft.as_dict()
        plp = eng.PLCalcParallel.from_flat(flat)
        o = plo.OracleProblem(flat)
        fp, P = eng.generate_param_order_vector(ft.free, lf=False)
        x0 = plp.set_freeparams(P, False, fp)
        o.set_freeparams(False)
        for B in (1, 32, 64):
            X = np.repeat(x0[:, None], B, axis=1)
            X[: ft.numnodes - 1, :] = 1.0 + 0.1 * np.arange(B)[None, :]
            plp.batch_configure(B)
            for mode in (eng.GRAD_ANALYTIC, eng.GRAD_AD):
                F, G = plp.calc_pl_grad_batch(X, mode=mode)
                for b in (0, B - 1):
                    xb = np.ascontiguousarray(X[:, b])
                    f, g = o.calc_pl_and_gradient(xb) if mode == eng.GRAD_ANALYTIC else o.calc_function_gradient(xb)
                    assert close_f(F[b], f) and rel_err(G[:, b], g) < G_TOL
        plp.close()


def test_gradient_computed_with_calc_pl_for_callback_pairs():
    """calc_pl -> calc_gradient pairs (every TNC / NLopt callback, optimize_tnc.cpp:75-105): from the second pair on calc_pl
    computes the analytic gradient in the same launch and calc_gradient returns it without another evaluation - the same
    numbers, bit for bit; calc_pl twice in a row (the annealer) switches that off; a state change between the two calls
    (smoothing, CV mask) forces a fresh evaluation."""
    gd = Golden("M1155")
    flat = gd.flat
    n = len(flat["parents"])
    fp, P = eng.generate_param_order_vector(flat["free"], lf=False)
    rng = np.random.RandomState(3)
    plp = eng.PLCalcParallel.from_flat(flat)
    x0 = plp.set_freeparams(P, False, fp)
    plp.smoothing = 10.0
    xs = []
    for _ in range(4):
        x = x0.copy()
        x[: n - 1] = 0.01 * np.exp(0.3 * rng.randn(n - 1))
        xs.append(x)
    ref = eng.PLCalcParallel.from_flat(flat)                 # never sees a pair twice: always the two-evaluation path
    ref.set_freeparams(P, False, fp)
    ref.smoothing = 10.0
    want = [ref.calc_pl_grad(x, eng.GRAD_ANALYTIC) for x in xs]
    ref.close()
    for k, x in enumerate(xs):                               # pair 0 arms the speculation, pairs 1.. use it
        f = plp.calc_pl(x)
        g = plp.calc_gradient(x)
        assert f == want[k][0] and np.array_equal(g, want[k][1])
    # the annealer's pattern: objective only; the next pair is again exact
    for x in xs[:3]:
        assert plp.calc_pl(x) == plp.calc_pl(x)
    f = plp.calc_pl(xs[3])
    assert f == want[3][0] and np.array_equal(plp.calc_gradient(xs[3]), want[3][1])
    # a state change between calc_pl and calc_gradient: the gradient of the NEW objective at the same point
    f = plp.calc_pl(xs[0]); plp.calc_gradient(xs[0])        # arm
    f = plp.calc_pl(xs[1])
    plp.smoothing = 1000.0
    g = plp.calc_gradient(xs[1])
    f2, g2 = plp.calc_pl_grad(xs[1], eng.GRAD_ANALYTIC)
    assert np.array_equal(g, g2) and not np.array_equal(g, want[1][1])
    plp.close()


def test_device_blocks_of_closed_engines_are_reused_and_can_be_released():
    """Work space of a closed engine stays with the library for the next engine of the process (tens of GB for a large CV
    batch: cudaMalloc / cudaFree of that size cost tenths of a second per engine); tpl_release_cached_memory returns it."""
    import torch
    from treepl_b200 import batchopt, cv
    gd = Golden("big_test")
    flat = gd.flat
    n = len(flat["parents"])
    fp, P = eng.generate_param_order_vector(flat["free"], lf=False)
    lo, hi = batchopt.reference_bounds(flat, fp, P)
    x0 = cv.start_vector(flat, P, n - 1)
    B = 256
    X0 = np.repeat(x0[:, None], B, axis=1)
    eng.release_cached_memory()
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]

    def run():
        plp = eng.PLCalcParallel.from_flat(flat)
        plp.set_freeparams(P, False, fp)
        plp.batch_configure(B)
        X, F, info = plp.optimize_batch(X0, lo, hi, method=1, max_steps=4)
        plp.close()
        return F

    F1 = run()
    held = free0 - torch.cuda.mem_get_info()[0]
    assert held > 300e6, held                                  # ~12 [P][B] arrays + the tree work planes are still held
    F2 = run()                                                 # the second engine runs in the first one's blocks
    assert np.array_equal(F1, F2)
    assert free0 - torch.cuda.mem_get_info()[0] <= held + 64e6
    eng.release_cached_memory()
    assert free0 - torch.cuda.mem_get_info()[0] < 64e6


@pytest.mark.parametrize("B", [130, 162, 200, 256])
def test_host_buffer_pipeline_handles_ragged_batches(B):
    """tpl_calc_pl_grad_batch with host buffers evaluates B >= 128 vectors as pipelined pieces of up to 64 (copies of piece k+1
    and k-1 overlap the kernels of piece k); a ragged tail is shared between the last two pieces so that none falls below the
    32 vectors the descriptor kernel needs.  Every vector must equal its evaluation in a small, unpipelined batch (gradient rows
    are written once by their owner: bitwise; the objective's partial sums are grouped per launch: to rounding) and the oracle."""
    gd = Golden("big_test")
    flat = gd.flat
    n = len(flat["parents"])
    fp, P = eng.generate_param_order_vector(flat["free"], lf=False)
    plp = eng.PLCalcParallel.from_flat(flat)
    x0 = plp.set_freeparams(P, False, fp)
    plp.smoothing = 3.0
    rng = np.random.RandomState(B)
    X = np.repeat(x0[:, None], B, axis=1)
    X[: n - 1, :] = 0.005 * np.exp(0.3 * rng.randn(n - 1, B))
    lam = 10.0 ** rng.uniform(-1, 2, B)
    plp.batch_configure(B, lane_smoothing=lam)
    F, G = plp.calc_pl_grad_batch(X, mode=eng.GRAD_AD)
    assert np.isfinite(F).all() and (F < 1e15).all()
    o = plo.OracleProblem(flat)
    o.set_freeparams(False)
    for lo_b in (0, B - 40):                                    # the head and the ragged tail
        sub = slice(lo_b, lo_b + 40)
        plp.batch_configure(40, lane_smoothing=lam[sub])
        Fs, Gs = plp.calc_pl_grad_batch(np.ascontiguousarray(X[:, sub]), mode=eng.GRAD_AD)
        assert np.allclose(F[sub], Fs, rtol=1e-13, atol=0)
        assert np.array_equal(G[:, sub], Gs)
    for b in (0, B - 1, B - 33):
        o.set_smoothing(lam[b])
        f, g = o.calc_function_gradient(np.ascontiguousarray(X[:, b]))
        assert close_f(F[b], f), (b, F[b], f)
        assert rel_err(G[:, b], g) < G_TOL
    plp.close()


def test_pinned_host_buffers_are_evaluated_in_place():
    """tpl_calc_pl_grad_batch on PINNED host buffers (tpl_host_alloc), 32 <= B < 128: the descriptor kernel reads X from and
    writes G to the host buffers through their device mapping (no staging copies).  Bitwise equal to the staged path on pageable
    buffers of the same batch."""
    import ctypes as C
    gd = Golden("big_test")
    flat = gd.flat
    n = len(flat["parents"])
    fp, P = eng.generate_param_order_vector(flat["free"], lf=False)
    plp = eng.PLCalcParallel.from_flat(flat)
    x0 = plp.set_freeparams(P, False, fp)
    plp.smoothing = 2.0
    B = 64
    rng = np.random.RandomState(11)
    X = np.repeat(x0[:, None], B, axis=1)
    X[: n - 1, :] = 0.005 * np.exp(0.3 * rng.randn(n - 1, B))
    plp.batch_configure(B)
    F0, G0 = plp.calc_pl_grad_batch(X, mode=eng.GRAD_AD)             # pageable numpy buffers: staged
    assert plp.last_kernel_generation() == 4
    L = plp._L
    nb = P * B * 8
    px, pg, pf = L.tpl_host_alloc(nb), L.tpl_host_alloc(nb), L.tpl_host_alloc(B * 8)
    assert px and pg and pf
    try:
        Xp = np.ctypeslib.as_array(C.cast(px, C.POINTER(C.c_double)), shape=(P, B))
        Gp = np.ctypeslib.as_array(C.cast(pg, C.POINTER(C.c_double)), shape=(P, B))
        Fp = np.ctypeslib.as_array(C.cast(pf, C.POINTER(C.c_double)), shape=(B,))
        Xp[:] = X
        Gp[:] = -7.0
        dp = C.POINTER(C.c_double)
        rc = L.tpl_calc_pl_grad_batch(plp._h, B, C.cast(px, dp), C.cast(pf, dp), C.cast(pg, dp), eng.GRAD_AD)
        assert rc >= 0, L.tpl_last_error()
        assert np.array_equal(Fp, F0)
        assert np.array_equal(Gp, G0)
        assert np.array_equal(Xp, X)                                 # inputs untouched
        plp.force_kernel(5)                                          # pinned buffers through the staged path
        Gp[:] = -7.0
        rc = L.tpl_calc_pl_grad_batch(plp._h, B, C.cast(px, dp), C.cast(pf, dp), C.cast(pg, dp), eng.GRAD_AD)
        assert rc >= 0
        assert np.array_equal(Fp, F0) and np.array_equal(Gp, G0)
    finally:
        L.tpl_host_free(px); L.tpl_host_free(pg); L.tpl_host_free(pf)
        plp.close()
