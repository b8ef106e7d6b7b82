"""TEST INFRASTRUCTURE ONLY: ctypes binding of the plain-C oracle (oracle/pl_oracle.c).

Same call surface as oracle/refbind.RefProblem so parity tests can run either behind one
interface.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libpl_oracle.so")

_lib = None


def build(force=False):
    src = os.path.join(_HERE, "pl_oracle.c")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return ORACLE_SO


class _Problem(C.Structure):
    _ip, _dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    _fields_ = [("n", C.c_int), ("parent", _ip), ("child_off", _ip), ("children", _ip), ("free_", _ip),
                ("c", _dp), ("lf", _dp), ("min_", _dp), ("max_", _dp), ("pen_min", _dp), ("pen_max", _dp),
                ("freeparams", _ip), ("cv", _ip), ("numparams", C.c_int), ("lf_mode", C.c_int),
                ("log_pen", C.c_int), ("smoothing", C.c_double), ("penalty_boundary", C.c_double),
                ("rates", _dp), ("dates", _dp), ("durations", _dp)]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(ORACLE_SO)
        pp, dp, ip = C.POINTER(_Problem), C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.plo_calc_pl.restype = C.c_double
        L.plo_calc_pl.argtypes = [pp, dp]
        L.plo_calc_gradient.restype = None
        L.plo_calc_gradient.argtypes = [pp, dp, dp]
        L.plo_calc_function_gradient.restype = C.c_double
        L.plo_calc_function_gradient.argtypes = [pp, dp, dp]
        L.plo_generate_param_order_vector.restype = C.c_int
        L.plo_generate_param_order_vector.argtypes = [C.c_int, ip, C.c_int, ip, ip]
        L.plo_set_freeparams_emit.restype = C.c_int
        L.plo_set_freeparams_emit.argtypes = [pp, dp]
        _lib = L
    return _lib


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class OracleProblem:
    """flat: dict with parents, child_counts, children, free, c, lf, min, max, pen_min, pen_max,
    start_dates, start_rates, start_durations (reference numbering)."""

    def __init__(self, flat, log_pen=False, smoothing=1.0, penalty_boundary=0.0025):
        i32 = lambda k: np.ascontiguousarray(flat[k], dtype=np.int32).copy()
        f64 = lambda k: np.ascontiguousarray(flat[k], dtype=np.float64).copy()
        self.a = {k: i32(k) for k in ("parents", "child_counts", "children", "free")}
        self.a.update({k: f64(k) for k in ("c", "lf", "min", "max", "pen_min", "pen_max")})
        n = self.a["parents"].shape[0]
        off = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(self.a["child_counts"], out=off[1:])
        self.a["child_off"] = off
        self.a["rates"] = f64("start_rates")
        self.a["dates"] = f64("start_dates")
        self.a["durations"] = f64("start_durations")
        self.a["cv"] = np.zeros(n, dtype=np.int32)
        self.a["freeparams"] = np.full(2 * n, -1, dtype=np.int32)
        self.p = _Problem()
        self.p.n = n
        for fld, key in (("parent", "parents"), ("child_off", "child_off"), ("children", "children"),
                         ("free_", "free"), ("freeparams", "freeparams"), ("cv", "cv")):
            setattr(self.p, fld, _ip(self.a[key]))
        for fld, key in (("c", "c"), ("lf", "lf"), ("min_", "min"), ("max_", "max"), ("pen_min", "pen_min"),
                         ("pen_max", "pen_max"), ("rates", "rates"), ("dates", "dates"), ("durations", "durations")):
            setattr(self.p, fld, _dp(self.a[key]))
        self.p.log_pen = int(log_pen)
        self.p.smoothing = float(smoothing)
        self.p.penalty_boundary = float(penalty_boundary)
        self.p.lf_mode = 0
        self.p.numparams = 0
        self._cv_set = False

    numnodes = property(lambda self: int(self.p.n))
    numparams = property(lambda self: int(self.p.numparams))

    def set_smoothing(self, s):
        self.p.smoothing = float(s)

    def set_log_pen(self, v):
        self.p.log_pen = int(v)

    def set_penalty_boundary(self, v):
        self.p.penalty_boundary = float(v)

    def set_state(self, rates, dates):
        self.a["rates"][:] = rates
        self.a["dates"][:] = dates

    def set_cv(self, cv):
        if cv is None:
            self.a["cv"][:] = 0
            self._cv_set = False
        else:
            self.a["cv"][:] = np.asarray(cv, dtype=np.int32)
            self._cv_set = True

    def set_freeparams(self, lf, use_cv=False):
        L = lib()
        cvp = _ip(self.a["cv"]) if (use_cv and self._cv_set) else None
        self.p.numparams = L.plo_generate_param_order_vector(self.p.n, _ip(self.a["free"]), int(lf), cvp,
                                                             _ip(self.a["freeparams"]))
        self.p.lf_mode = int(lf)
        x0 = np.zeros(self.p.numparams, dtype=np.float64)
        k = L.plo_set_freeparams_emit(C.byref(self.p), _dp(x0))
        return x0[:k]

    def ints(self, name):
        return self.a[name].copy()

    def doubles(self, name):
        return self.a[name].copy()

    def calc_pl(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.p.numparams
        return lib().plo_calc_pl(C.byref(self.p), _dp(x))

    def calc_pl_and_gradient(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        g = np.zeros_like(x)
        f = lib().plo_calc_pl(C.byref(self.p), _dp(x))
        lib().plo_calc_gradient(C.byref(self.p), _dp(x), _dp(g))
        return f, g

    def calc_function_gradient(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        g = np.full_like(x, np.nan)
        f = lib().plo_calc_function_gradient(C.byref(self.p), _dp(x), _dp(g))
        return f, g
