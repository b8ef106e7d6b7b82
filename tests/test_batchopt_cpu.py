"""Host logic of the batched optimiser (tests/torch_lbfgs_proto.py, an independent eager-torch L-BFGS) on CPU, with the ORACLE as the evaluator,
against the reference's own END-TO-END results (`seed=1, nthreads=1` runs of the reference):
the optimum it finds is the reference's final PL value and dated tree."""
import numpy as np
import pytest
import torch

from golden_util import Golden
from oracle import plo
import torch_lbfgs_proto
from treepl_b200 import batchopt, cv
from treepl_b200 import treeflat as tf


def _fit(name, lams, max_iter):
    gd = Golden(name)
    flat = gd.flat
    n = len(flat["parents"])
    fp, P = tf.generate_param_order_vector(flat["free"], lf=False)
    lo, hi = batchopt.reference_bounds(flat, fp, P)
    B = len(lams)
    probs = [plo.OracleProblem(flat) for _ in lams]
    for o, lam in zip(probs, lams):
        o.set_smoothing(lam)
        o.set_freeparams(False)

    def evaluator(X):
        Xn = X.numpy()
        F = np.zeros(B)
        G = np.zeros((P, B))
        for b, o in enumerate(probs):
            f, g = o.calc_function_gradient(np.ascontiguousarray(Xn[:, b]))
            F[b] = f
            G[:, b] = np.nan_to_num(g) if f < 1e15 else 0.0
        return torch.from_numpy(F), torch.from_numpy(G)

    x0 = cv.start_vector(flat, P, n - 1)
    par = flat["parents"]
    ds = float(np.median(np.maximum(flat["start_dates"][par[1:]] - flat["start_dates"][1:], 1e-12)))
    opt = torch_lbfgs_proto.BatchedLBFGS(evaluator, P, B, n - 1, lo, hi, date_scale=ds)
    X, F, info = opt.minimize(np.repeat(x0[:, None], B, axis=1), max_iter=max_iter)
    return flat, fp, X, F, info


def test_test_cppr8s_final_pl_and_dated_tree():
    """examples/test.cppr8s (smooth = 0.1): reference final PL 234.23049 and dated tree
    (((a:3.997337,b:3.997337):4.002084,((c:3.207058,d:3.207058):1.600094,e:4.807152):3.192270):2.000579,(f:6.626589,g:6.626589):3.373411)"""
    flat, fp, X, F, info = _fit("test", [0.1, 0.1], 600)
    assert info["converged"].all()
    assert abs(F[0] - 234.23049) <= 1e-4 * 234.23049           # the reference stops a little short of the optimum
    assert F[0] <= 234.23049
    n = 13
    dates = {i: X[fp[n + i], 0] for i in range(n) if fp[n + i] != -1}
    want = {1: 6.626589, 4: 10 - 2.000579, 5: 4.807152, 7: 3.207058, 10: 3.997337}
    for node, h in want.items():
        assert abs(dates[node] - h) <= 2e-3 * h, (node, dates[node], h)


def test_pl4203_and_m1155_final_pl():
    """final PL of examples/pl4203.cppr8s at its CV-selected smoothing (100): 390.04656;
    of examples/M1155.cppr8s at smoothing 10: 252.3558"""
    flat, fp, X, F, info = _fit("pl4203", [100.0, 1.0], 2500)
    assert abs(F[0] - 390.04656) <= 2e-6 * 390.04656, F[0]
    flat, fp, X, F, info = _fit("M1155", [10.0, 10.0], 4000)
    assert abs(F[0] - 252.3558) <= 2e-6 * 252.3558, F[0]
