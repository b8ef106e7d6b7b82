"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/libtreepl_ref.so.

The library is the UNMODIFIED reference (every source of its Makefile.in except main.cpp,
compiled from /root/reference/src by oracle/Makefile against NLopt 2.4.2 and ADOL-C 2.6.3
built from the reference's own deps/ tarballs) behind the thin harness oracle/ref_harness.cpp.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the product
package treepl_b200/ never does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libtreepl_ref.so")

_lib = None


def available():
    return os.path.exists(REF_SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_SO)
        vp, ci, cd = C.c_void_p, C.c_int, C.c_double
        ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
        L.ref_from_newick.restype = vp
        L.ref_from_newick.argtypes = [C.c_char_p, ci, ci, ci, ci, C.POINTER(C.c_char_p), ip, dp, ip, dp, ci, ci]
        L.ref_from_flat.restype = vp
        L.ref_from_flat.argtypes = [ci, ip, ip, ip, ip] + [dp] * 9 + [ci]
        L.ref_free.argtypes = [vp]
        L.ref_numnodes.argtypes = [vp]
        L.ref_get_ints.argtypes = [vp, ci, ip]
        L.ref_get_doubles.argtypes = [vp, ci, dp]
        L.ref_get_name.argtypes = [vp, ci, C.c_char_p, ci]
        L.ref_set_smoothing.argtypes = [vp, cd]
        L.ref_set_log_pen.argtypes = [vp, ci]
        L.ref_set_penalty_boundary.argtypes = [vp, cd]
        L.ref_set_state.argtypes = [vp, dp, dp]
        L.ref_set_cv.argtypes = [vp, ip]
        L.ref_set_freeparams.argtypes = [vp, ci, ci, dp]
        L.ref_calc_pl.restype = cd
        L.ref_calc_pl.argtypes = [vp, dp]
        L.ref_calc_pl_and_gradient.restype = cd
        L.ref_calc_pl_and_gradient.argtypes = [vp, dp, dp]
        L.ref_calc_function_gradient.restype = cd
        L.ref_calc_function_gradient.argtypes = [vp, dp, dp]
        L.ref_calc_function_gradient_adolc.restype = cd
        L.ref_calc_function_gradient_adolc.argtypes = [vp, dp, dp]
        L.ref_time_evals.restype = cd
        L.ref_time_evals.argtypes = [vp, dp, ci, ci, ci]
        L.ref_cv.restype = ci
        L.ref_cv.argtypes = [vp, ci, dp, ci, ip, ip, ci, ci, dp, dp, dp]
        L.ref_hw_threads.restype = ci
        L.ref_multistart.restype = cd
        L.ref_multistart.argtypes = [vp, ci, dp, ci, ci, dp]
        L.ref_optimize.restype = ci
        L.ref_optimize.argtypes = [vp, ci, ci, cd, cd, ci, dp, dp, C.c_char_p, ci]
        _lib = L
    return _lib


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


_INTS = {"parents": 0, "child_counts": 1, "free": 2, "children": 3, "freeparams": 4}
_DBLS = {"c": 0, "lf": 1, "min": 2, "max": 3, "pen_min": 4, "pen_max": 5, "start_dates": 6,
         "start_rates": 7, "start_durations": 8, "rates": 9, "dates": 10, "durations": 11}


class RefProblem:
    """One reference `pl_calc_parallel` plus the flat vectors it borrows."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("reference harness could not build the problem")
        self.h = C.c_void_p(handle)
        self.numparams = 0

    @classmethod
    def from_newick(cls, newick, numsites, calibrations=(), seed=1, log_pen=False, collapse=False, set1=False):
        """calibrations: iterable of (list_of_tip_names, min_or_None, max_or_None)."""
        L = lib()
        n = len(calibrations)
        tips = (C.c_char_p * max(n, 1))(*[(" ".join(c[0])).encode() for c in calibrations] or [b""])
        hmin = np.array([c[1] is not None for c in calibrations] or [0], dtype=np.int32)
        hmax = np.array([c[2] is not None for c in calibrations] or [0], dtype=np.int32)
        vmin = np.array([c[1] if c[1] is not None else 0.0 for c in calibrations] or [0.0], dtype=np.float64)
        vmax = np.array([c[2] if c[2] is not None else 0.0 for c in calibrations] or [0.0], dtype=np.float64)
        h = L.ref_from_newick(newick.strip().encode(), int(numsites), int(collapse), int(set1), n, tips,
                              _ip(hmin), _dp(vmin), _ip(hmax), _dp(vmax), int(seed), int(log_pen))
        return cls(h)

    @classmethod
    def from_flat(cls, flat, log_pen=False):
        """flat: dict with the keys of FlatTree.as_dict() (see treepl_b200/treeflat.py)."""
        L = lib()
        i32 = lambda k: np.ascontiguousarray(flat[k], dtype=np.int32)
        f64 = lambda k: np.ascontiguousarray(flat[k], dtype=np.float64)
        keep = [i32("parents"), i32("child_counts"), i32("children"), i32("free")] + \
               [f64(k) for k in ("c", "lf", "min", "max", "pen_min", "pen_max",
                                 "start_dates", "start_rates", "start_durations")]
        h = L.ref_from_flat(len(keep[0]), *[_ip(a) for a in keep[:4]], *[_dp(a) for a in keep[4:]], int(log_pen))
        return cls(h)

    def close(self):
        if self.h:
            lib().ref_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def numnodes(self):
        return lib().ref_numnodes(self.h)

    def ints(self, name):
        n = lib().ref_get_ints(self.h, _INTS[name], None)
        out = np.zeros(n, dtype=np.int32)
        lib().ref_get_ints(self.h, _INTS[name], _ip(out))
        return out

    def doubles(self, name):
        n = lib().ref_get_doubles(self.h, _DBLS[name], None)
        out = np.zeros(n, dtype=np.float64)
        lib().ref_get_doubles(self.h, _DBLS[name], _dp(out))
        return out

    def name(self, node):
        buf = C.create_string_buffer(256)
        lib().ref_get_name(self.h, node, buf, 256)
        return buf.value.decode()

    def flat(self):
        d = {k: self.ints(k) for k in ("parents", "child_counts", "free", "children")}
        d.update({k: self.doubles(k) for k in ("c", "lf", "min", "max", "pen_min", "pen_max",
                                               "start_dates", "start_rates", "start_durations")})
        return d

    def set_smoothing(self, s):
        lib().ref_set_smoothing(self.h, float(s))

    def set_log_pen(self, v):
        lib().ref_set_log_pen(self.h, int(v))

    def set_penalty_boundary(self, v):
        lib().ref_set_penalty_boundary(self.h, float(v))

    def set_state(self, rates, dates):
        r = np.ascontiguousarray(rates, dtype=np.float64)
        d = np.ascontiguousarray(dates, dtype=np.float64)
        lib().ref_set_state(self.h, _dp(r), _dp(d))

    def set_cv(self, cv):
        if cv is None:
            lib().ref_set_cv(self.h, None)
        else:
            a = np.ascontiguousarray(cv, dtype=np.int32)
            lib().ref_set_cv(self.h, _ip(a))

    def set_freeparams(self, lf, use_cv=False):
        """-> x0 (numpy).  Also records numparams."""
        n = self.numnodes
        buf = np.zeros(3 * n + 2, dtype=np.float64)
        self.numparams = lib().ref_set_freeparams(self.h, int(lf), int(use_cv), _dp(buf))
        return buf[: self.numparams].copy()

    def calc_pl(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.numparams
        return lib().ref_calc_pl(self.h, _dp(x))

    def calc_pl_and_gradient(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.numparams
        g = np.zeros_like(x)
        f = lib().ref_calc_pl_and_gradient(self.h, _dp(x), _dp(g))
        return f, g

    def calc_function_gradient(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.numparams
        g = np.zeros_like(x)
        f = lib().ref_calc_function_gradient(self.h, _dp(x), _dp(g))
        return f, g

    def calc_function_gradient_adolc(self, x):
        """ADOL-C tape (calc_pl_function_gradient_adolc / calc_lf_function_gradient_adolc)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.numparams
        g = np.zeros_like(x)
        f = lib().ref_calc_function_gradient_adolc(self.h, _dp(x), _dp(g))
        return f, g

    def cross_validate(self, smoothing_grid, groups, randomcv=False, nthreads=1):
        """The reference's Langley-Fitch fit + cross-validation loop (main.cpp:497-822) around its own
        optimize_full (annealer + TNC).  groups: list of lists of tip node numbers.  Call right after
        from_newick(seed=...) for the libc rand() sequence of a reference run.
        -> dict(chisq [K], full_pl [K], seconds {lf, full, folds})."""
        lam = np.ascontiguousarray(smoothing_grid, dtype=np.float64)
        off = np.zeros(len(groups) + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(g) for g in groups])
        nodes = np.ascontiguousarray([v for g in groups for v in g], dtype=np.int32)
        chisq = np.zeros(len(lam))
        full = np.zeros(len(lam))
        secs = np.zeros(3)
        rc = lib().ref_cv(self.h, len(lam), _dp(lam), len(groups), _ip(off), _ip(nodes), int(bool(randomcv)),
                          int(nthreads), _dp(chisq), _dp(full), _dp(secs))
        if rc != 0:
            raise RuntimeError("reference cross-validation failed to initialise (rc=%d)" % rc)
        return {"chisq": chisq, "full_pl": full, "seconds": {"lf": secs[0], "full": secs[1], "folds": secs[2]}}

    def optimize(self, which, x0, numiter=10000, ftol=1e-7, xtol=1e-7, seed=1):
        """The reference's own optimiser adapter `which` (codes: oracle/ref_harness.cpp::ref_optimize) from x0
        -> (rc, x, f, log)."""
        x = np.ascontiguousarray(x0, dtype=np.float64).copy()
        assert x.size == self.numparams
        f = C.c_double()
        log = C.create_string_buffer(1 << 16)
        rc = lib().ref_optimize(self.h, int(which), int(numiter), float(ftol), float(xtol), int(seed), _dp(x), C.byref(f), log, len(log))
        return rc, x, f.value, log.value.decode(errors="replace")

    def multistart(self, X0, nthreads=1, seed=1):
        """X0 [nstarts, numparams]: the reference's optimize_full from every start, one engine per OpenMP thread
        -> (objectives [nstarts], seconds)."""
        X0 = np.ascontiguousarray(X0, dtype=np.float64)
        assert X0.ndim == 2 and X0.shape[1] == self.numparams
        f = np.zeros(X0.shape[0])
        secs = lib().ref_multistart(self.h, X0.shape[0], _dp(X0), int(nthreads), int(seed), _dp(f))
        return f, secs

    def time_evals(self, x, reps, nthreads, with_grad=True):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return lib().ref_time_evals(self.h, _dp(x), int(reps), int(nthreads), int(with_grad))


def hw_threads():
    return lib().ref_hw_threads()
