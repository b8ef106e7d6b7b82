"""Truncated-Newton optimiser logic and the analytic Hessian-vector product on the CPU
(treepl_b200/csrc/tpl_tn_core.h compiled with g++ behind tests/tn_sim.cpp; the ORACLE supplies objective and gradient).
  * H v  == central finite differences of the oracle's gradient (z-space, with CV masks and per-vector smoothing);
  * diag == e_i^T H e_i;
  * the optimiser reaches the reference's END-TO-END results."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from golden_util import Golden
from oracle import plo
from treepl_b200 import batchopt, cv
from treepl_b200 import treeflat as tf
from test_lbfgs_sim import transformed_bounds, EVAL_FN

HERE = os.path.dirname(os.path.abspath(__file__))
SIM_SO = os.path.join(HERE, "_build", "libtn_sim.so")
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
up = C.POINTER(C.c_ubyte)


def load_sim():
    src = os.path.join(HERE, "tn_sim.cpp")
    deps = [src] + [os.path.join(HERE, "..", "treepl_b200", "csrc", f) for f in ("tpl_tn_core.h", "tpl_lbfgs_core.h")]
    os.makedirs(os.path.dirname(SIM_SO), exist_ok=True)
    if not os.path.exists(SIM_SO) or os.path.getmtime(SIM_SO) < max(os.path.getmtime(d) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SIM_SO, src, "-lm"])
    lib = C.CDLL(SIM_SO)
    lib.tn_sim_hessvec.restype = None
    lib.tn_sim_hessvec.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, ip, ip, ip, dp, dp, dp, dp, dp, up, dp, dp, C.c_double, dp, dp, dp, dp, dp, C.c_int]
    lib.tn_sim_treesolve.restype = None
    lib.tn_sim_treesolve.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, ip, ip, ip, dp, dp, dp, dp, dp, up, dp, dp, C.c_double, dp, dp, dp, dp, dp, C.c_int]
    lib.tn_sim.restype = C.c_int
    lib.tn_sim.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                           C.c_double, C.c_double, EVAL_FN, C.c_void_p, ip, ip, ip,
                           C.c_int, ip, ip, ip, ip, ip, dp, dp, dp, dp, dp, up, dp, C.c_double, C.c_int, C.c_int]
    return lib


@pytest.fixture(scope="module")
def sim():
    return load_sim()


def _d(a):
    return a.ctypes.data_as(dp)


def _i(a):
    return a.ctypes.data_as(ip)


class Problem:
    """flat arrays in the form the kernels take them + oracle evaluators (one per vector)"""

    def __init__(self, flat, lams, masks=None, log_pen=False):
        self.flat = flat
        self.log_pen = bool(log_pen)
        n = self.n = len(flat["parents"])
        self.fp, self.P = tf.generate_param_order_vector(flat["free"], lf=False)
        self.nr = n - 1
        self.B = len(lams)
        self.lams = np.ascontiguousarray(lams, dtype=np.float64)
        self.parent = np.ascontiguousarray(flat["parents"], dtype=np.int32)
        self.child_off = np.zeros(n + 1, dtype=np.int32)
        self.child_off[1:] = np.cumsum(flat["child_counts"])
        self.children = np.ascontiguousarray(flat["children"], dtype=np.int32)
        self.rslot = np.ascontiguousarray(self.fp[:n], dtype=np.int32)
        self.dslot = np.ascontiguousarray(self.fp[n:], dtype=np.int32)
        self.c = np.ascontiguousarray(flat["c"], dtype=np.float64)
        self.hfix = np.ascontiguousarray(flat["start_dates"], dtype=np.float64)
        self.rfix = np.ascontiguousarray(flat["start_rates"], dtype=np.float64)
        self.pmin = np.ascontiguousarray(flat["pen_min"], dtype=np.float64)
        self.pmax = np.ascontiguousarray(flat["pen_max"], dtype=np.float64)
        self.mask = None
        self.probs = []
        for b, lam in enumerate(lams):
            o = plo.OracleProblem(flat)
            o.set_smoothing(lam)
            o.set_log_pen(self.log_pen)
            if masks is not None and masks[b]:
                cvm = np.zeros(n, np.int32)
                cvm[list(masks[b])] = 1
                o.set_cv(cvm)
            o.set_freeparams(False)          # FULL layout: a masked node keeps its (ignored) rate row, as in the batched kernels
            self.probs.append(o)
        if masks is not None:
            self.mask = np.zeros((self.B, n), dtype=np.uint8)
            for b, m in enumerate(masks):
                self.mask[b, list(m)] = 1
        par = flat["parents"]
        ds = float(np.median(np.maximum(flat["start_dates"][par[1:]] - flat["start_dates"][1:], 1e-12)))
        self.sc = np.ones(self.P)
        self.sc[self.nr:] = ds
        lower, upper = batchopt.reference_bounds(flat, self.fp, self.P)
        lo, hi = transformed_bounds(lower, upper, self.nr)
        self.lo, self.hi = lo / self.sc, hi / self.sc
        self.nev = 0

    def tree_args(self):
        return [self.n, _i(self.parent), _i(self.child_off), _i(self.children), _i(self.rslot), _i(self.dslot), _d(self.c), _d(self.hfix),
                _d(self.rfix), _d(self.pmin), _d(self.pmax), self.mask.ctypes.data_as(up) if self.mask is not None else None]

    def eval(self, X):
        """X [P,B] original parameters -> F [B], G [P,B]; a masked node's rate row gets gradient 0"""
        F = np.zeros(self.B)
        G = np.zeros((self.P, self.B))
        self.nev += 1
        for b, o in enumerate(self.probs):
            f, g = o.calc_function_gradient(np.ascontiguousarray(X[:, b]))
            F[b] = f
            if f < 1e15:
                G[:, b] = g
                if self.mask is not None:
                    for i in np.flatnonzero(self.mask[b]):
                        G[self.rslot[i], b] = 0.0
        return F, G

    def callback(self):
        def fn(Xp, Fp, Gp, ctx):
            X = np.ctypeslib.as_array(Xp, shape=(self.P, self.B))
            F, G = self.eval(X)
            np.ctypeslib.as_array(Fp, shape=(self.B,))[:] = F
            np.ctypeslib.as_array(Gp, shape=(self.P, self.B))[:] = G
        return EVAL_FN(fn)

    def to_x(self, Z):
        X = Z * self.sc[:, None]
        X[: self.nr] = np.exp(X[: self.nr])
        return X

    def grad_z(self, Z):
        X = self.to_x(Z)
        F, G = self.eval(X)
        Gz = G * self.sc[:, None]
        Gz[: self.nr] *= X[: self.nr]
        return F, Gz, X, G


def _masked_layout_ok(flat):
    return True


@pytest.mark.parametrize("name,lams,masked,log_pen", [("test", [0.1, 10.0], False, False), ("M1155", [10.0, 0.1, 1000.0], True, False),
                                                      ("pl4203", [1.0], True, False), ("vib", [10.0, 1000.0], True, True),
                                                      ("test", [0.1, 100.0], False, True)])
def test_hessian_vector_product_matches_finite_differences(sim, name, lams, masked, log_pen):
    flat = Golden(name).flat
    n = len(flat["parents"])
    tips = np.flatnonzero(flat["child_counts"] == 0)
    masks = None
    if masked:
        masks = [[int(tips[3 + 2 * b])] for b in range(len(lams))]
        masks[0] = []
    pr = Problem(flat, lams, masks, log_pen=log_pen)
    P, B = pr.P, pr.B
    rng = np.random.RandomState(5)
    x0 = cv.start_vector(flat, P, n - 1)
    X0 = np.repeat(x0[:, None], B, axis=1)
    X0[: n - 1] *= np.exp(0.2 * rng.randn(n - 1, B))
    Z = X0 / pr.sc[:, None]
    Z[: n - 1] = np.log(X0[: n - 1]) / pr.sc[: n - 1, None]
    V = rng.randn(P, B)
    if masks is not None:
        for b, m in enumerate(masks):
            for i in m:
                V[pr.rslot[i], b] = 0.0          # a masked rate row is not a variable
    F, Gz, X, G = pr.grad_z(Z)
    W = np.zeros((P, B))
    DG = np.zeros((P, B))
    Xc, Gc, Vc = np.ascontiguousarray(X), np.ascontiguousarray(G), np.ascontiguousarray(V)
    sim.tn_sim_hessvec(P, B, *pr.tree_args(), _d(pr.sc), _d(pr.lams), 0.0025, _d(Xc), _d(Gc), _d(Vc), _d(W), _d(DG), int(pr.log_pen))
    eps = 1e-6
    _, g1, _, _ = pr.grad_z(Z + eps * V)
    _, g2, _, _ = pr.grad_z(Z - eps * V)
    fd = (g1 - g2) / (2 * eps)
    for b in range(B):
        err = np.abs(W[:, b] - fd[:, b]).max() / np.abs(fd[:, b]).max()
        assert err < 5e-8, (name, b, err)
    # diagonal: e_i^T H e_i for a sample of rows
    for row in list(rng.choice(P, 12, replace=False)):
        E = np.zeros((P, B))
        E[row] = 1.0
        Wd = np.zeros((P, B))
        D2 = np.zeros((P, B))
        sim.tn_sim_hessvec(P, B, *pr.tree_args(), _d(pr.sc), _d(pr.lams), 0.0025, _d(Xc), _d(Gc), _d(np.ascontiguousarray(E)), _d(Wd), _d(D2), int(pr.log_pen))
        assert np.allclose(Wd[row], DG[row], rtol=1e-12, atol=1e-300), (row, Wd[row], DG[row])


def fit(sim, name, lams, max_steps=20000, max_cg=500, masks=None, ftol=1e-13, patience=3, cg_per_step=4, precond=0, log_pen=False,
        flat=None):
    flat = Golden(name).flat if flat is None else flat
    pr = Problem(flat, lams, masks, log_pen=log_pen)
    n = pr.n
    x0 = cv.start_vector(flat, pr.P, n - 1)
    X = np.ascontiguousarray(np.repeat(x0[:, None], pr.B, axis=1))
    F = np.zeros(pr.B)
    status = np.zeros(pr.B, np.int32)
    iters = np.zeros(pr.B, np.int32)
    ncg = np.zeros(pr.B, np.int32)
    cb = pr.callback()
    steps = sim.tn_sim(pr.P, pr.B, pr.nr, _d(X), _d(F), _d(pr.lo), _d(pr.hi), _d(pr.sc), max_steps, max_cg, cg_per_step, patience, 40, ftol, 1e-8, 1e-4, 0.9,
                       cb, None, _i(status), _i(iters), _i(ncg), *pr.tree_args(), _d(pr.lams), 0.0025, int(precond), int(log_pen))
    return pr, X, F, status, iters, ncg, steps


def test_tn_reproduces_reference_final_pl_values(sim):
    pr, X, F, status, iters, ncg, steps = fit(sim, "test", [0.1, 0.1])
    assert ((status >= 1) & (status <= 2)).all(), status
    assert abs(F[0] - 234.23049) <= 1e-4 * 234.23049 and F[0] <= 234.23049 and F[0] == F[1]
    pr, X, F, status, iters, ncg, steps = fit(sim, "pl4203", [100.0, 1.0])
    assert ((status >= 1) & (status <= 2)).all(), status
    assert abs(F[0] - 390.04656) <= 2e-6 * 390.04656, F
    pr, X, F, status, iters, ncg, steps = fit(sim, "M1155", [1000.0, 100.0, 10.0, 1.0, 0.1])
    assert ((status >= 1) & (status <= 2)).all(), status
    want = [434.8684339, 337.8462252, 252.3558021, 223.2864918, 219.3542532]      # converged L-BFGS / reference final PL 252.3558
    assert np.allclose(F, want, rtol=2e-9), F
    assert steps < 1500, steps          # far fewer evaluations than L-BFGS needs (3,000+ for smoothing 0.1 alone)


def _dense_hessian(sim, pr, X, G):
    P, B = pr.P, pr.B
    H = np.zeros((B, P, P))
    for row in range(P):
        E = np.zeros((P, B))
        E[row] = 1.0
        W, D = np.zeros((P, B)), np.zeros((P, B))
        sim.tn_sim_hessvec(P, B, *pr.tree_args(), _d(pr.sc), _d(pr.lams), 0.0025, _d(X), _d(G), _d(np.ascontiguousarray(E)), _d(W), _d(D), int(pr.log_pen))
        H[:, :, row] = W.T
    return H


@pytest.mark.parametrize("name,lams,log_pen", [("test", [0.1, 100.0], False), ("pl4203", [100.0], False), ("test", [1.0], True)])
def test_tree_factorisation_is_an_exact_solve_of_the_tree_part_of_the_hessian(sim, name, lams, log_pen):
    """M = H_z minus the sibling couplings of the root's children (the only non-tree entries).  Where M is positive
    definite the level sweeps are an exact solve (dense reference), frozen rows come out as 0, and M^-1 H_z has all
    eigenvalues 1 except the few that the dropped couplings move — what CG then needs a handful of iterations for."""
    flat = Golden(name).flat
    n = len(flat["parents"])
    pr = Problem(flat, lams, None, log_pen=log_pen)
    P, B = pr.P, pr.B
    # a point near the optimum, where H_z is positive definite
    _, Xopt, F, status, _, _, _ = fit(sim, name, lams, precond=1, cg_per_step=2, log_pen=log_pen)
    rng = np.random.RandomState(1)
    X0 = Xopt * np.exp(0.01 * rng.randn(P, B))
    Z = X0 / pr.sc[:, None]
    Z[: n - 1] = np.log(X0[: n - 1]) / pr.sc[: n - 1, None]
    F, Gz, X, G = pr.grad_z(Z)
    X, G = np.ascontiguousarray(X), np.ascontiguousarray(G)
    H = _dense_hessian(sim, pr, X, G)
    rootkids = [i for i in range(n) if flat["parents"][i] == 0]
    rhs = np.ascontiguousarray(rng.randn(P, B))
    free = np.ones((P, B))
    frozen = rng.choice(P, 3, replace=False)
    free[frozen] = 0.0
    out = np.zeros((P, B))
    sim.tn_sim_treesolve(P, B, *pr.tree_args(), _d(pr.sc), _d(pr.lams), 0.0025, _d(X), _d(G), _d(free), _d(rhs), _d(out), int(log_pen))
    keep = np.setdiff1d(np.arange(P), frozen)
    for b in range(B):
        M = H[b].copy()
        for a in rootkids:
            for c in rootkids:
                if a != c:
                    M[pr.rslot[a], pr.rslot[c]] = 0.0
        Mk = M[np.ix_(keep, keep)]
        assert np.linalg.eigvalsh(Mk).min() > 0
        want = np.linalg.solve(Mk, rhs[keep, b])
        assert np.abs(out[keep, b] - want).max() <= 1e-9 * np.abs(want).max()
        assert np.all(out[frozen, b] == 0.0)
        ev = np.linalg.eigvals(np.linalg.solve(Mk, H[b][np.ix_(keep, keep)])).real
        assert (np.abs(ev - 1.0) > 1e-6).sum() <= 2 * len(rootkids)


def test_tree_preconditioned_tn_converges_where_jacobi_does_not(sim):
    """pl4203 at smoothing 1e-4 and the clock tree at smoothing 1e4 (rates ~ 1e3: the smoothing curvature exceeds the
    likelihood curvature by 1e6 - 1e8): the Jacobi-preconditioned solves crawl, the tree-preconditioned ones take one or
    two CG iterations per Newton step; both reach the same optimum where both converge."""
    pr, X, F, status, iters, ncg, steps = fit(sim, "pl4203", [100.0, 1e-4], precond=1, cg_per_step=2, max_steps=3000)
    assert ((status >= 1) & (status <= 2)).all() and ncg.sum() <= 2 * iters.sum() + 8, (status, iters, ncg)
    assert abs(F[0] - 390.04656) <= 2e-6 * 390.04656 and abs(F[1] - 214.56107) <= 2e-6 * 214.56107, F   # e2e_reference.json after_opt_calc
    pj, Xj, Fj, sj, itj, ncgj, stepsj = fit(sim, "clock", [1e4, 1.0], precond=0, max_steps=3000)
    pt, Xt, Ft, st, itt, ncgt, stepst = fit(sim, "clock", [1e4, 1.0], precond=1, cg_per_step=2, max_steps=3000)
    assert ((st >= 1) & (st <= 2)).all()
    assert np.all(np.abs(Ft - Fj) <= 1e-8 * np.abs(Fj)), (Ft, Fj)
    assert ncgt.sum() * 10 < ncgj.sum(), (ncgt, ncgj)


def test_tn_with_log_penalty_matches_reference_objective_semantics(sim):
    """log_pen (pl_calc_parallel.cpp:866-879) in the truncated-Newton Hessian: the optimum of the oracle's log-penalty objective"""
    pr, X, F, status, iters, ncg, steps = fit(sim, "vib", [1000.0, 10.0], precond=1, cg_per_step=2, log_pen=True)
    assert (status == 1).all()
    f, g = pr.eval(X)
    assert np.all(np.abs(f - F) <= 1e-12 * np.abs(F))
    gz = g.copy()
    gz[: pr.nr] *= X[: pr.nr]
    assert np.all(np.abs(gz).max(axis=0) <= 1e-6 * np.abs(F))
