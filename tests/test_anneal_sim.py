"""Batched annealer logic (treepl_b200/csrc/tpl_anneal_core.h) on the CPU: the O(degree) delta evaluation keeps the running
objective EQUAL to a full evaluation by the oracle's calc_pl (pl_calc_parallel.cpp:73-130) after every batch of proposals, with
CV masks, per-chain smoothing, calibration barriers and the log penalty; the chain follows the reference's rules
(src/siman_calc_par.cpp:80-229): never worse than its start, discards infeasible proposals, cools every 100 iterations."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from golden_util import Golden
from oracle import plo
from treepl_b200 import cv
from treepl_b200 import treeflat as tf

HERE = os.path.dirname(os.path.abspath(__file__))
SIM_SO = os.path.join(HERE, "_build", "libanneal_sim.so")
dp, ip, up = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_ubyte)


def load_sim():
    src = os.path.join(HERE, "anneal_sim.cpp")
    deps = [src] + [os.path.join(HERE, "..", "treepl_b200", "csrc", f) for f in ("tpl_anneal_core.h", "tpl_lbfgs_core.h")]
    os.makedirs(os.path.dirname(SIM_SO), exist_ok=True)
    if not os.path.exists(SIM_SO) or os.path.getmtime(SIM_SO) < max(os.path.getmtime(d) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SIM_SO, src, "-lm"])
    lib = C.CDLL(SIM_SO)
    lib.sa_sim_create.restype = C.c_void_p
    lib.sa_sim_create.argtypes = [C.c_int] * 4 + [ip] * 6 + [C.c_int] + [dp] * 6 + [up, C.c_int, ip, dp, dp, dp, C.c_double, C.c_int, dp] + \
        [C.c_double] * 4 + [C.c_int, C.c_double, C.c_ulonglong, ip]
    lib.sa_sim_steps.restype = C.c_int
    lib.sa_sim_steps.argtypes = [C.c_void_p, C.c_int]
    lib.sa_sim_state.argtypes = [C.c_void_p, C.c_int, dp]
    lib.sa_sim_free.argtypes = [C.c_void_p]
    return lib


@pytest.fixture(scope="module")
def sim():
    return load_sim()


class AnnealProblem:
    def __init__(self, flat, lams, masks=None, log_pen=False):
        self.flat = flat
        n = self.n = len(flat["parents"])
        self.fp, self.P = tf.generate_param_order_vector(flat["free"], lf=False)
        self.nr = n - 1
        self.B = len(lams)
        self.lams = np.ascontiguousarray(lams, dtype=np.float64)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self.parent = i32(flat["parents"])
        self.child_off = np.zeros(n + 1, dtype=np.int32)
        self.child_off[1:] = np.cumsum(flat["child_counts"])
        self.children = i32(flat["children"])
        self.rslot, self.dslot = i32(self.fp[:n]), i32(self.fp[n:])
        self.date_node = i32([i for i in range(n) if self.fp[n + i] != -1])
        self.c, self.lf = f64(flat["c"]), f64(flat["lf"])
        self.hfix, self.rfix = f64(flat["start_dates"]), f64(flat["start_rates"])
        self.hmin, self.hmax = f64(flat["min"]), f64(flat["max"])
        bar = [i for i in range(n) if flat["pen_min"][i] != -1 or flat["pen_max"][i] != -1]
        self.bar_node = i32(bar or [0])
        self.nbar = len(bar)
        self.bar_pmin, self.bar_pmax = f64(flat["pen_min"][bar] if bar else [-1.0]), f64(flat["pen_max"][bar] if bar else [-1.0])
        self.mask = None
        self.log_pen = log_pen
        self.probs = []
        for b, lam in enumerate(lams):
            o = plo.OracleProblem(flat, log_pen=log_pen)
            o.set_smoothing(lam)
            if masks is not None and masks[b]:
                cvm = np.zeros(n, np.int32)
                cvm[list(masks[b])] = 1
                o.set_cv(cvm)
            o.set_freeparams(False)
            self.probs.append(o)
        if masks is not None:
            self.mask = np.zeros((self.B, n), dtype=np.uint8)
            for b, m in enumerate(masks):
                self.mask[b, list(m)] = 1

    def create(self, sim, X, temp=50.0, cool=0.999, rt=0.07, dt=0.25, maxit=15000, stop=0.001, seed=1):
        self.ok = np.zeros(self.B, dtype=np.int32)
        d = lambda a: a.ctypes.data_as(dp)
        i = lambda a: a.ctypes.data_as(ip)
        return sim.sa_sim_create(self.P, self.B, self.nr, self.n, i(self.parent), i(self.child_off), i(self.children), i(self.rslot), i(self.dslot),
                                 i(self.date_node), len(self.date_node), d(self.c), d(self.lf), d(self.hfix), d(self.rfix), d(self.hmin), d(self.hmax),
                                 self.mask.ctypes.data_as(up) if self.mask is not None else None, self.nbar, i(self.bar_node), d(self.bar_pmin),
                                 d(self.bar_pmax), d(self.lams), 0.0025, int(self.log_pen), d(X), temp, cool, rt, dt, maxit, stop, seed, i(self.ok))

    def full_objective(self, X):
        return np.array([o.calc_pl(np.ascontiguousarray(X[:, b])) for b, o in enumerate(self.probs)])


def state(sim, h, b):
    out = np.zeros(10)
    sim.sa_sim_state(h, b, out.ctypes.data_as(dp))
    return dict(zip(("f", "temperature", "step_rt", "step_dt", "iteration", "discarded", "rate_accepts", "rate_trials", "date_accepts",
                     "date_trials"), out))


@pytest.mark.parametrize("name,lams,masked,log_pen", [("test", [0.1, 10.0], False, False), ("test_two_cal", [1.0, 0.1], False, True),
                                                       ("M1155", [10.0, 0.1, 1000.0], True, False), ("pl4203", [1.0, 100.0], True, True)])
def test_running_objective_equals_full_evaluation(sim, name, lams, masked, log_pen):
    flat = Golden(name).flat
    n = len(flat["parents"])
    tips = np.flatnonzero(flat["child_counts"] == 0)
    masks = None
    if masked:
        masks = [[int(tips[3 + 2 * b])] for b in range(len(lams))]
        masks[0] = []
    pr = AnnealProblem(flat, lams, masks, log_pen)
    x0 = cv.start_vector(flat, pr.P, n - 1)
    X = np.ascontiguousarray(np.repeat(x0[:, None], pr.B, axis=1))
    X[: n - 1] *= np.exp(0.1 * np.random.RandomState(2).randn(n - 1, pr.B))
    f0 = pr.full_objective(X)
    h = pr.create(sim, X, seed=7)
    assert pr.ok.all()
    for b in range(pr.B):
        assert abs(state(sim, h, b)["f"] - f0[b]) <= 1e-12 * abs(f0[b])
    last = f0.copy()
    for chunk in range(12):
        sim.sa_sim_steps(h, 250)
        full = pr.full_objective(X)                      # X is updated in place by the chains
        for b in range(pr.B):
            st = state(sim, h, b)
            assert abs(st["f"] - full[b]) <= 1e-10 * abs(full[b]), (chunk, b, st["f"], full[b])
            assert st["f"] <= last[b] + 1e-9 * abs(last[b])          # exp((-cur-test)/T) ~ 0: only improvements are kept
            last[b] = st["f"]
    st = state(sim, h, 0)
    assert st["iteration"] + st["discarded"] == 3000 and st["iteration"] > 1000
    assert abs(st["temperature"] - 50.0 * 0.999 ** (st["iteration"] // 100)) < 1e-9
    assert last[0] < f0[0]
    sim.sa_sim_free(h)


def test_chain_stops_on_temperature_or_iteration_cap(sim):
    flat = Golden("test").flat
    pr = AnnealProblem(flat, [0.1, 0.1])
    x0 = cv.start_vector(flat, pr.P, pr.n - 1)
    X = np.ascontiguousarray(np.repeat(x0[:, None], 2, axis=1))
    h = pr.create(sim, X, maxit=700, seed=3)
    while sim.sa_sim_steps(h, 500) > 0:
        pass
    assert state(sim, h, 0)["iteration"] == 701                      # `iteration > maxit` (:224)
    sim.sa_sim_free(h)
    X = np.ascontiguousarray(np.repeat(x0[:, None], 2, axis=1))
    h = pr.create(sim, X, temp=1.0, cool=0.5, stop=0.1, maxit=15000, seed=3)
    while sim.sa_sim_steps(h, 500) > 0:
        pass
    assert state(sim, h, 1)["iteration"] == 400                      # 1.0 * 0.5^4 = 0.0625 <= 0.1 after the 4th cooling
    sim.sa_sim_free(h)
