"""Logic of the device-resident optimiser (treepl_b200/csrc/tpl_lbfgs_core.h: per-vector state machine, Gram-matrix
two-loop recursion) on the CPU.  tests/lbfgs_sim.cpp compiles the very functions the CUDA kernels call and replaces
the thread grid by loops; the ORACLE evaluates the objective.  Checked against the reference's own END-TO-END
results (`seed=1, nthreads=1` runs): final PL values and the dated tree."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from golden_util import Golden
from oracle import plo
from treepl_b200 import batchopt, cv
from treepl_b200 import treeflat as tf

HERE = os.path.dirname(os.path.abspath(__file__))
SIM_SO = os.path.join(HERE, "_build", "liblbfgs_sim.so")
EVAL_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)


def load_sim():
    src = os.path.join(HERE, "lbfgs_sim.cpp")
    core = os.path.join(HERE, "..", "treepl_b200", "csrc", "tpl_lbfgs_core.h")
    os.makedirs(os.path.dirname(SIM_SO), exist_ok=True)
    if not os.path.exists(SIM_SO) or os.path.getmtime(SIM_SO) < max(os.path.getmtime(src), os.path.getmtime(core)):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SIM_SO, src, "-lm"])
    lib = C.CDLL(SIM_SO)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.lbfgs_sim.restype = C.c_int
    lib.lbfgs_sim.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, C.c_int, C.c_int, C.c_int,
                              C.c_double, C.c_double, C.c_double, EVAL_FN, C.c_void_p, ip, ip, C.c_int, ip, ip, dp, C.c_double]
    return lib


@pytest.fixture(scope="module")
def sim():
    return load_sim()


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def transformed_bounds(lower, upper, nr):
    """the bound transform of tpl_optimize_batch (tpl_engine.cu), restated: bounds in w = ln r | h"""
    lo = np.array(lower, dtype=np.float64)
    hi = np.array(upper, dtype=np.float64)
    lo[:nr] = np.log(np.maximum(lo[:nr], 1e-300))
    hi[:nr] = np.log(hi[:nr])
    dlo, dhi = lo[nr:], hi[nr:]
    margin = 1e-9 * np.maximum(1.0, np.abs(np.where(dhi < 1e19, dhi, dlo)))
    dlo += margin
    dhi -= np.where(dhi < 1e19, margin, 0.0)
    return lo, hi


def diagonal_scales(flat, fp, X, lams):
    """the per-element scales of opt_scale_kernel (tpl_lbfgs.cuh), restated in numpy: sc = 1 / sqrt(H_ww) with
    H_ww = r d + 2 lambda r^2 (1 + #children) for w = ln r, and the sum of c / d^2 over the adjacent branches for w = h"""
    n = len(flat["parents"])
    par = flat["parents"]
    P, B = X.shape
    SC = np.ones((P, B))
    c = flat["c"]
    nch = flat["child_counts"]
    for b in range(B):
        dates = flat["start_dates"].astype(np.float64).copy()
        rates = np.zeros(n)
        for i in range(n):
            if fp[n + i] != -1:
                dates[i] = X[fp[n + i], b]
            if fp[i] != -1:
                rates[i] = X[fp[i], b]
        dur = np.maximum(dates[par[1:]] - dates[1:], 2.2250738585072014e-308)
        hd = np.zeros(n)
        w = c[1:] / dur ** 2
        np.add.at(hd, np.arange(1, n), w)
        np.add.at(hd, par[1:], w)
        for i in range(1, n):
            if fp[i] != -1:
                r = rates[i]
                SC[fp[i], b] = 1.0 / np.sqrt(max(r * dur[i - 1] + 2.0 * lams[b] * r * r * (1 + nch[i]), 1e-300))
        for i in range(n):
            if fp[n + i] != -1:
                SC[fp[n + i], b] = 1.0 / np.sqrt(max(hd[i], 1e-300))
    return SC


def fit(sim, name, lams, max_steps, history=8, boundary_frac=0.9, ftol=1e-12, patience=5, scaling="global", flat=None, x0=None):
    flat = Golden(name).flat if flat is None else flat
    n = len(flat["parents"])
    fp, P = tf.generate_param_order_vector(flat["free"], lf=False)
    lower, upper = batchopt.reference_bounds(flat, fp, P)
    B = len(lams)
    probs = [plo.OracleProblem(flat) for _ in lams]
    for o, lam in zip(probs, lams):
        o.set_smoothing(lam)
        o.set_freeparams(False)
    nev = [0]

    def evaluator(Xp, Fp, Gp, ctx):
        X = np.ctypeslib.as_array(Xp, shape=(P, B))
        F = np.ctypeslib.as_array(Fp, shape=(B,))
        G = np.ctypeslib.as_array(Gp, shape=(P, B))
        nev[0] += 1
        for b, o in enumerate(probs):
            f, g = o.calc_function_gradient(np.ascontiguousarray(X[:, b]))
            F[b] = f
            if f < 1e15:
                G[:, b] = g

    x0 = cv.start_vector(flat, P, n - 1) if x0 is None else x0
    par = flat["parents"]
    lo, hi = transformed_bounds(lower, upper, n - 1)
    X = np.ascontiguousarray(np.repeat(x0[:, None], B, axis=1)) if x0.ndim == 1 else np.ascontiguousarray(x0)
    if scaling == "diag":
        SC = np.ascontiguousarray(diagonal_scales(flat, fp, X, lams))
    else:
        ds = float(np.median(np.maximum(flat["start_dates"][par[1:]] - flat["start_dates"][1:], 1e-12)))
        SC = np.ones((P, B))
        SC[n - 1:] = ds
    F = np.zeros(B)
    status = np.zeros(B, dtype=np.int32)
    iters = np.zeros(B, dtype=np.int32)
    cb = EVAL_FN(evaluator)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    parent = np.ascontiguousarray(flat["parents"], dtype=np.int32)
    dslot = np.ascontiguousarray(fp[n:], dtype=np.int32)
    hfix = np.ascontiguousarray(flat["start_dates"], dtype=np.float64)
    steps = sim.lbfgs_sim(history, P, B, n - 1, _dp(X), _dp(F), _dp(lo), _dp(hi), _dp(SC), max_steps, patience, 40, ftol, 1e-8, 1e-4, cb, None,
                          ip(status), ip(iters), n, ip(parent), ip(dslot), _dp(hfix), boundary_frac)
    assert steps == nev[0]
    return flat, fp, X, F, status, iters, steps


def test_sim_test_cppr8s_final_pl_and_dated_tree(sim):
    """examples/test.cppr8s (smooth = 0.1): reference final PL 234.23049 and its dated tree"""
    flat, fp, X, F, status, iters, steps = fit(sim, "test", [0.1, 0.1], 3000)
    assert (status > 0).all() and (status < 4).all(), status
    assert abs(F[0] - 234.23049) <= 1e-4 * 234.23049
    assert F[0] <= 234.23049
    assert F[0] == F[1]                                        # identical vectors take identical paths
    n = 13
    dates = {i: X[fp[n + i], 0] for i in range(n) if fp[n + i] != -1}
    want = {1: 6.626589, 4: 10 - 2.000579, 5: 4.807152, 7: 3.207058, 10: 3.997337}
    for node, h in want.items():
        assert abs(dates[node] - h) <= 2e-3 * h, (node, dates[node], h)


@pytest.mark.parametrize("history", [4, 8, 16])
def test_sim_final_pl_of_example_configs(sim, history):
    """final PL of examples/pl4203.cppr8s at smoothing 100: 390.04656; of examples/M1155.cppr8s at smoothing 10: 252.3558.
    Vectors with different smoothing values advance independently (their line searches de-synchronise)."""
    flat, fp, X, F, status, iters, steps = fit(sim, "pl4203", [100.0, 1.0], 6000, history)
    assert (status > 0).all() and (status < 4).all(), status
    assert abs(F[0] - 390.04656) <= 2e-6 * 390.04656, F[0]
    flat, fp, X, F, status, iters, steps = fit(sim, "M1155", [10.0, 1000.0], 8000, history)
    assert (status > 0).all() and (status < 4).all(), status
    assert abs(F[0] - 252.3558) <= 2e-6 * 252.3558, F[0]
