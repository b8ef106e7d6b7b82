"""Host-side mirror of the reference's `pl_calc_parallel` (src/pl_calc_parallel.h:17-82) over the
C ABI of libtreepl_b200.so (include/treepl_b200.h).

Same method names, argument meaning and error behaviour as the reference class, so parity tests
read like calls on the reference object:

    plp = PLCalcParallel()
    plp.setup_starting_bits(parents, child_counts, free, char_durations, log_fact, vmin, vmax,
                            start_dates, start_rates, start_durations,
                            children_vec=..., pen_min=..., pen_max=...)     # main.cpp:483-490
    plp.set_log_pen(False); plp.smoothing = 0.1
    n = generate_param_order_vector(freeparams_out, lf, cvnodes, free)      # utils.cpp:295
    x0 = plp.set_freeparams(numparams, lf, freeparams)                      # pl_calc_parallel.cpp:920
    f = plp.calc_pl(x); g = plp.calc_gradient(x)                            # function_plcp, optimize_tnc.cpp:75

There is NO CPU fallback: if the CUDA library is missing or no device is usable, construction
raises.  PyTorch is not imported here; device pointers for the batched entry point are plain ints.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

LARGE = 1e15
GRAD_ANALYTIC = 0
GRAD_AD = 1

_lib = None


class OptOptions(C.Structure):
    """tpl_opt_options (include/treepl_b200.h); 0 = library default"""
    _fields_ = [("history", C.c_int), ("max_steps", C.c_int), ("patience", C.c_int), ("max_ls", C.c_int),
                ("check_every", C.c_int), ("use_graph", C.c_int), ("method", C.c_int), ("max_cg", C.c_int), ("cg_per_step", C.c_int), ("precond", C.c_int), ("ftol", C.c_double), ("gtol", C.c_double),
                ("armijo", C.c_double), ("date_scale", C.c_double), ("boundary_frac", C.c_double), ("tree_wide_min", C.c_int), ("tree_coop", C.c_int), ("hv_pipe", C.c_int), ("reserved", C.c_int)]


class AnnealOptions(C.Structure):
    """tpl_anneal_options (include/treepl_b200.h); 0 = the reference's default"""
    _fields_ = [("temperature", C.c_double), ("cool", C.c_double), ("rate_step", C.c_double), ("date_step", C.c_double),
                ("stop_temperature", C.c_double), ("max_iterations", C.c_int), ("seed", C.c_ulonglong)]


class OptResult(C.Structure):
    _fields_ = [("steps", C.c_int), ("n_done", C.c_int), ("launches", C.c_int), ("cg_iterations", C.c_int), ("device_ms", C.c_double),
                ("mean_log10_step", C.c_double)]


class EngineError(RuntimeError):
    pass


def release_cached_memory():
    """tpl_release_cached_memory: give the device memory kept from closed engines back to the driver"""
    load_library().tpl_release_cached_memory()


def lib_path():
    return _build.LIB


def load_library():
    """Load libtreepl_b200.so (built in-tree by treepl_b200/build.py).  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise EngineError("libtreepl_b200.so is not built (python -m treepl_b200.build); "
                          "the PL engine has no CPU fallback")
    L = C.CDLL(path)
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.tpl_create.restype = ci
    L.tpl_create.argtypes = [C.POINTER(vp), ci, ci, ip, ip, ip, ip] + [dp] * 9
    L.tpl_destroy.restype = None
    L.tpl_destroy.argtypes = [vp]
    L.tpl_last_error.restype = C.c_char_p
    for name, args in (("tpl_set_smoothing", [vp, cd]), ("tpl_set_log_pen", [vp, ci]),
                       ("tpl_set_penalty_boundary", [vp, cd]), ("tpl_set_cv_nodes", [vp, ip]),
                       ("tpl_set_state", [vp, dp, dp]),
                       ("tpl_generate_param_order_vector", [ci, ip, ci, ip, ip]),
                       ("tpl_set_freeparams", [vp, ci, ci, ip, dp]), ("tpl_numnodes", [vp]), ("tpl_numparams", [vp]),
                       ("tpl_calc_gradient", [vp, dp, dp]), ("tpl_get_rates", [vp, dp]), ("tpl_get_dates", [vp, dp]),
                       ("tpl_get_durations", [vp, dp]), ("tpl_batch_configure", [vp, ci, dp, ip, ip]),
                       ("tpl_calc_pl_grad_batch", [vp, ci, dp, dp, dp, ci]),
                       ("tpl_calc_pl_grad_batch_dev", [vp, ci, vp, vp, vp, ci, vp]),
                       ("tpl_last_launch_count", [vp]), ("tpl_device_sm", [vp]),
                       ("tpl_debug_math", [ci, dp, dp, dp]), ("tpl_set_timing", [vp, ci]),
                       ("tpl_get_timing", [vp, dp, dp]), ("tpl_debug_force_kernel", [vp, ci]),
                       ("tpl_last_kernel_generation", [vp])):
        fn = getattr(L, name)
        fn.restype = ci
        fn.argtypes = args
    for name, args in (("tpl_calc_pl", [vp, dp]), ("tpl_calc_function_gradient", [vp, dp, dp]),
                       ("tpl_calc_pl_grad", [vp, dp, dp, ci])):
        fn = getattr(L, name)
        fn.restype = cd
        fn.argtypes = args
    L.tpl_optimize_batch.restype = ci
    L.tpl_optimize_batch.argtypes = [vp, ci, dp, dp, dp, dp, ci, C.POINTER(OptOptions), ip, ip, C.POINTER(OptResult)]
    L.tpl_anneal_batch.restype = ci
    L.tpl_anneal_batch.argtypes = [vp, ci, dp, dp, C.POINTER(AnnealOptions), ip, ip]
    L.tpl_optimize_batch_cols.restype = ci
    L.tpl_optimize_batch_cols.argtypes = [vp, ci, ci, dp, ip, dp, dp, dp, ci, C.POINTER(OptOptions), ip, ip, C.POINTER(OptResult), ci, ip, dp, dp]
    L.tpl_debug_hessvec.restype = ci
    L.tpl_debug_hessvec.argtypes = [vp, ci, dp, dp, dp, dp, dp]
    L.tpl_debug_treesolve.restype = ci
    L.tpl_debug_treesolve.argtypes = [vp, ci, dp, dp, dp, dp, dp, dp, dp]
    L.tpl_host_alloc.restype = vp
    L.tpl_host_alloc.argtypes = [C.c_ulonglong]
    L.tpl_host_free.restype = None
    L.tpl_host_free.argtypes = [vp]
    _lib = L
    return L


EXPORTED_SYMBOLS = (
    "tpl_create", "tpl_destroy", "tpl_last_error", "tpl_set_smoothing", "tpl_set_log_pen",
    "tpl_set_penalty_boundary", "tpl_set_cv_nodes", "tpl_set_state", "tpl_generate_param_order_vector",
    "tpl_set_freeparams", "tpl_numnodes", "tpl_numparams", "tpl_calc_pl", "tpl_calc_gradient",
    "tpl_calc_function_gradient", "tpl_calc_pl_grad", "tpl_get_rates", "tpl_get_dates", "tpl_get_durations",
    "tpl_batch_configure", "tpl_calc_pl_grad_batch", "tpl_calc_pl_grad_batch_dev", "tpl_last_launch_count",
    "tpl_host_alloc", "tpl_host_free", "tpl_device_sm", "tpl_debug_math", "tpl_set_timing", "tpl_get_timing", "tpl_debug_force_kernel",
    "tpl_last_kernel_generation", "tpl_release_cached_memory", "tpl_optimize_batch", "tpl_optimize_batch_cols", "tpl_debug_hessvec", "tpl_debug_treesolve", "tpl_anneal_batch",
)


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def generate_param_order_vector(free, lf, cvnodes=None):
    """utils.cpp:295-340 through the library's host helper -> (freeparams int32[2N], numparams)."""
    L = load_library()
    free = _i32(free)
    n = free.shape[0]
    fp = np.empty(2 * n, dtype=np.int32)
    cv = _i32(cvnodes) if cvnodes is not None else None
    nump = L.tpl_generate_param_order_vector(n, _ip(free), int(bool(lf)), _ip(cv) if cv is not None else None, _ip(fp))
    if nump < 0:
        raise EngineError(L.tpl_last_error().decode())
    return fp, nump


def debug_math(x):
    """-> (log, reciprocal) as the kernels compute them (test hook)."""
    L = load_library()
    x = _f64(x)
    lo, rc = np.zeros_like(x), np.zeros_like(x)
    if L.tpl_debug_math(x.shape[0], _dp(x), _dp(lo), _dp(rc)) < 0:
        raise EngineError(L.tpl_last_error().decode())
    return lo, rc


class PLCalcParallel:
    """B200 stand-in for `class pl_calc_parallel`."""

    def __init__(self, device=-1):
        self._L = load_library()
        self._h = C.c_void_p()
        self._device = device
        self.numnodes = 0
        self.numparams = 0
        self.lf = False
        self._smoothing = 1.0
        self.freeparams = None

    # -- lifetime ---------------------------------------------------------------------------------
    def setup_starting_bits(self, parents_nds, child_cnts, vfree, char_dur, log_fact_ch, vmn, vmx,
                            start_dates, start_rates, start_durations, children_vec=None,
                            pen_min=None, pen_max=None):
        """pl_calc_parallel.cpp:34-71 + the member assignments of main.cpp:488-490.
        children_vec: list of per-node child lists, or the flattened int array."""
        self.close()
        parents = _i32(parents_nds)
        n = parents.shape[0]
        if children_vec is None:
            raise ValueError("children_vec is required (main.cpp:490)")
        if len(children_vec) == n and n > 1 and not np.isscalar(children_vec[0]):
            flat = [c for kids in children_vec for c in kids]
            children = _i32(flat if flat else [0])
        else:
            children = _i32(children_vec)
        pmn = _f64(pen_min) if pen_min is not None else np.full(n, -1.0)
        pmx = _f64(pen_max) if pen_max is not None else np.full(n, -1.0)
        args = [_i32(child_cnts), children, _i32(vfree)]
        dbl = [_f64(char_dur), _f64(log_fact_ch), _f64(vmn), _f64(vmx), pmn, pmx, _f64(start_dates),
               _f64(start_rates), _f64(start_durations)]
        rc = self._L.tpl_create(C.byref(self._h), int(self._device), n, _ip(parents), *[_ip(a) for a in args],
                                *[_dp(a) for a in dbl])
        self._check(rc)
        self.numnodes = n
        self.free = _i32(vfree)
        self._smoothing = 1.0
        return self

    @classmethod
    def from_flat(cls, flat, device=-1):
        """flat: dict / FlatTree.as_dict() with the reference's flat vectors."""
        return cls(device).setup_starting_bits(
            flat["parents"], flat["child_counts"], flat["free"], flat["c"], flat["lf"], flat["min"], flat["max"],
            flat["start_dates"], flat["start_rates"], flat["start_durations"],
            children_vec=flat["children"], pen_min=flat["pen_min"], pen_max=flat["pen_max"])

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.tpl_destroy(self._h)
            self._h = C.c_void_p()

    delete_ad_arrays = close      # main.cpp:798

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise EngineError(self._L.tpl_last_error().decode())
        return rc

    # -- configuration ----------------------------------------------------------------------------
    @property
    def smoothing(self):
        return self._smoothing

    @smoothing.setter
    def smoothing(self, v):
        self._smoothing = float(v)
        self._check(self._L.tpl_set_smoothing(self._h, float(v)))

    def set_log_pen(self, v):
        self._check(self._L.tpl_set_log_pen(self._h, int(bool(v))))

    def set_penalty_boundary(self, v):
        self._check(self._L.tpl_set_penalty_boundary(self._h, float(v)))

    def set_cv_nodes(self, cvnodes):
        if cvnodes is None:
            self._check(self._L.tpl_set_cv_nodes(self._h, None))
        else:
            cv = _i32(cvnodes)
            assert cv.shape[0] == self.numnodes
            self._check(self._L.tpl_set_cv_nodes(self._h, _ip(cv)))

    def set_state(self, rates, dates):
        r, d = _f64(rates), _f64(dates)
        self._check(self._L.tpl_set_state(self._h, _dp(r), _dp(d)))

    def set_freeparams(self, nump, lf, freep):
        """-> initial parameter vector (the reference fills `params`)."""
        fp = _i32(freep)
        assert fp.shape[0] == 2 * self.numnodes
        out = np.zeros(int(nump), dtype=np.float64)
        k = self._check(self._L.tpl_set_freeparams(self._h, int(nump), int(bool(lf)), _ip(fp), _dp(out)))
        self.numparams = int(nump)
        self.lf = bool(lf)
        self.freeparams = fp.copy()
        return out[:k]

    # -- evaluation -------------------------------------------------------------------------------
    def _x(self, x):
        x = _f64(x)
        if x.shape[0] != self.numparams:
            raise ValueError("parameter vector has %d entries, engine expects %d" % (x.shape[0], self.numparams))
        return x

    def _ret(self, f):
        if f != f:      # NaN is how the C ABI reports an engine failure
            msg = self._L.tpl_last_error().decode()
            if msg:
                raise EngineError(msg)
        return f

    def calc_pl(self, params):
        x = self._x(params)
        return self._ret(self._L.tpl_calc_pl(self._h, _dp(x)))

    def calc_gradient(self, params):
        x = self._x(params)
        g = np.zeros_like(x)
        self._check(self._L.tpl_calc_gradient(self._h, _dp(x), _dp(g)))
        return g

    def calc_function_gradient(self, params):
        """-> (f, g); g is NaN-filled where the reference would have left it untouched."""
        x = self._x(params)
        g = np.full_like(x, np.nan)
        f = self._ret(self._L.tpl_calc_function_gradient(self._h, _dp(x), _dp(g)))
        return f, g

    calc_pl_function_gradient_adolc = calc_function_gradient      # bit-identical expression

    def calc_pl_grad(self, params, mode=GRAD_ANALYTIC, want_grad=True):
        x = self._x(params)
        g = np.full_like(x, np.nan) if want_grad else None
        f = self._ret(self._L.tpl_calc_pl_grad(self._h, _dp(x), _dp(g) if want_grad else None, int(mode)))
        return f, g

    @property
    def rates(self):
        out = np.zeros(self.numnodes)
        self._check(self._L.tpl_get_rates(self._h, _dp(out)))
        return out

    @property
    def dates(self):
        out = np.zeros(self.numnodes)
        self._check(self._L.tpl_get_dates(self._h, _dp(out)))
        return out

    @property
    def durations(self):
        out = np.zeros(self.numnodes)
        self._check(self._L.tpl_get_durations(self._h, _dp(out)))
        return out

    # -- batched ----------------------------------------------------------------------------------
    def batch_configure(self, B, lane_smoothing=None, masks=None):
        """masks: list of B iterables of masked node numbers (CV folds), or None."""
        lam = _f64(lane_smoothing) if lane_smoothing is not None else None
        off = nodes = None
        if masks is not None:
            assert len(masks) == B
            off = np.zeros(B + 1, dtype=np.int32)
            off[1:] = np.cumsum([len(m) for m in masks])
            nodes = _i32([v for m in masks for v in m] or [0])
        self._check(self._L.tpl_batch_configure(self._h, int(B), _dp(lam) if lam is not None else None,
                                                _ip(off) if off is not None else None,
                                                _ip(nodes) if nodes is not None else None))

    def calc_pl_grad_batch(self, X, mode=GRAD_ANALYTIC, want_grad=True):
        """X: array [P, B] (batch innermost, C order) -> (F[B], G[P, B] or None)."""
        X = _f64(X)
        P, B = X.shape
        if P != self.numparams:
            raise ValueError("X has %d rows, engine expects %d" % (P, self.numparams))
        F = np.zeros(B)
        G = np.zeros((P, B)) if want_grad else None
        self._check(self._L.tpl_calc_pl_grad_batch(self._h, B, _dp(X), _dp(F), _dp(G) if want_grad else None, int(mode)))
        return F, G

    def calc_pl_grad_batch_dev(self, B, dX, dF, dG, mode=GRAD_ANALYTIC, stream=0):
        """Device pointers (ints), asynchronous on `stream` (int handle, 0 = engine stream)."""
        self._check(self._L.tpl_calc_pl_grad_batch_dev(self._h, int(B), C.c_void_p(dX), C.c_void_p(dF),
                                                       C.c_void_p(dG) if dG else None, int(mode),
                                                       C.c_void_p(stream) if stream else None))

    def optimize_batch(self, X0, lower, upper, mode=GRAD_AD, **options):
        """Device-resident batched L-BFGS (tpl_optimize_batch): X0 [P, B] feasible start vectors ->
        (X [P, B], F [B], info).  options: fields of tpl_opt_options (history, max_steps, patience, max_ls,
        check_every, use_graph, method (0 L-BFGS, 1 truncated Newton), max_cg, cg_per_step, precond, ftol, gtol, armijo, date_scale, boundary_frac,
        tree_wide_min, tree_coop, hv_pipe)."""
        X = _f64(X0).copy()
        P, B = X.shape
        if P != self.numparams:
            raise ValueError("X0 has %d rows, engine expects %d" % (P, self.numparams))
        lo, hi = _f64(lower), _f64(upper)
        assert lo.shape[0] == P and hi.shape[0] == P
        F = np.zeros(B)
        status = np.zeros(B, dtype=np.int32)
        iters = np.zeros(B, dtype=np.int32)
        o = OptOptions()
        for k, v in options.items():
            if not hasattr(o, k):
                raise TypeError("unknown optimiser option %r" % k)
            setattr(o, k, v)
        res = OptResult()
        self._check(self._L.tpl_optimize_batch(self._h, B, _dp(X), _dp(F), _dp(lo), _dp(hi), int(mode), C.byref(o),
                                               _ip(status), _ip(iters), C.byref(res)))
        return X, F, {"steps": res.steps, "n_done": res.n_done, "launches": res.launches, "device_ms": res.device_ms,
                      "mean_log10_step": res.mean_log10_step, "cg_iterations": res.cg_iterations, "status": status, "iterations": iters}

    def optimize_batch_cols(self, Xstart, start_col, lower, upper, probe_rows=None, want_x=False, mode=GRAD_AD, **options):
        """tpl_optimize_batch_cols: vector b starts from column start_col[b] of Xstart [P, K]; probe_rows [nprobe, B] (X row
        per vector, -1 = none) selects what returns -> (F [B], probes [nprobe, B], X [P, B] or None, info)."""
        Xs = _f64(Xstart)
        P, K = Xs.shape
        if P != self.numparams:
            raise ValueError("Xstart has %d rows, engine expects %d" % (P, self.numparams))
        col = _i32(start_col)
        B = col.shape[0]
        lo, hi = _f64(lower), _f64(upper)
        assert lo.shape[0] == P and hi.shape[0] == P
        rows = _i32(probe_rows) if probe_rows is not None else np.zeros((0, B), dtype=np.int32)
        assert rows.ndim == 2 and rows.shape[1] == B
        nprobe = rows.shape[0]
        probes = np.zeros((nprobe, B))
        X = np.zeros((P, B)) if want_x else None
        F = np.zeros(B)
        status = np.zeros(B, dtype=np.int32)
        iters = np.zeros(B, dtype=np.int32)
        o = OptOptions()
        options.setdefault("method", 1)
        for k, v in options.items():
            if not hasattr(o, k):
                raise TypeError("unknown optimiser option %r" % k)
            setattr(o, k, v)
        res = OptResult()
        self._check(self._L.tpl_optimize_batch_cols(self._h, B, K, _dp(Xs), _ip(col), _dp(F), _dp(lo), _dp(hi), int(mode), C.byref(o),
                                                    _ip(status), _ip(iters), C.byref(res), nprobe, _ip(rows) if nprobe else None,
                                                    _dp(probes) if nprobe else None, _dp(X) if want_x else None))
        return F, probes, X, {"steps": res.steps, "n_done": res.n_done, "launches": res.launches, "device_ms": res.device_ms,
                              "mean_log10_step": res.mean_log10_step, "cg_iterations": res.cg_iterations, "status": status,
                              "iterations": iters}

    def anneal_batch(self, X0, **options):
        """Batched simulated annealing (tpl_anneal_batch): X0 [P, B] feasible starts -> (X, F, status, iterations).
        options: fields of tpl_anneal_options."""
        X = _f64(X0).copy()
        P, B = X.shape
        if P != self.numparams:
            raise ValueError("X0 has %d rows, engine expects %d" % (P, self.numparams))
        F = np.zeros(B)
        status = np.zeros(B, dtype=np.int32)
        iters = np.zeros(B, dtype=np.int32)
        o = AnnealOptions()
        for k, v in options.items():
            if not hasattr(o, k):
                raise TypeError("unknown annealer option %r" % k)
            setattr(o, k, v)
        self._check(self._L.tpl_anneal_batch(self._h, B, _dp(X), _dp(F), C.byref(o), _ip(status), _ip(iters)))
        return X, F, status, iters

    def debug_treesolve(self, X, free, rhs, sc_rows, rhs2=None):
        """-> M^-1 rhs through the block-tree factorisation kernels of the truncated-Newton method (test hook);
        with rhs2 also M^-1 rhs2 through the solve-only sweeps (same factors): -> (out, out2)."""
        X, Fr, R, sc = _f64(X), _f64(free), _f64(rhs), _f64(sc_rows)
        P, B = X.shape
        out = np.zeros((P, B))
        if rhs2 is None:
            self._check(self._L.tpl_debug_treesolve(self._h, B, _dp(X), _dp(Fr), _dp(R), _dp(sc), _dp(out), None, None))
            return out
        R2 = _f64(rhs2)
        out2 = np.zeros((P, B))
        self._check(self._L.tpl_debug_treesolve(self._h, B, _dp(X), _dp(Fr), _dp(R), _dp(sc), _dp(out), _dp(R2), _dp(out2)))
        return out, out2

    def debug_hessvec(self, X, V, sc_rows):
        """-> (H_z V, diag H_z) at X through the truncated-Newton kernels (test hook)."""
        X, V, sc = _f64(X), _f64(V), _f64(sc_rows)
        P, B = X.shape
        W, DG = np.zeros((P, B)), np.zeros((P, B))
        self._check(self._L.tpl_debug_hessvec(self._h, B, _dp(X), _dp(V), _dp(sc), _dp(W), _dp(DG)))
        return W, DG

    def set_timing(self, on):
        self._check(self._L.tpl_set_timing(self._h, int(bool(on))))

    def get_timing(self):
        """-> (evaluations, main-kernel ms, finalize-kernel ms) since the last call."""
        m, f = C.c_double(), C.c_double()
        cnt = self._check(self._L.tpl_get_timing(self._h, C.byref(m), C.byref(f)))
        return cnt, m.value, f.value

    def force_kernel(self, generation):
        self._check(self._L.tpl_debug_force_kernel(self._h, int(generation)))

    def last_kernel_generation(self):
        return self._L.tpl_last_kernel_generation(self._h)

    def last_launch_count(self):
        return self._L.tpl_last_launch_count(self._h)

    def device_sm(self):
        return self._L.tpl_device_sm(self._h)
