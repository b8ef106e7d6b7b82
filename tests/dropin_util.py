"""TEST INFRASTRUCTURE: loader for the drop-in builds of tests/dropin/ (the reference's own optimiser adapters and
CLI compiled UNMODIFIED against the engine's C++ mirror).  `make -C tests/dropin` needs /root/reference; on the GPU box
the prebuilt files under tests/_build/ travel with the snapshot."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
FLAT_ORDER_I = ("parents", "child_counts", "children", "free")
FLAT_ORDER_D = ("c", "lf", "min", "max", "pen_min", "pen_max", "start_dates", "start_rates", "start_durations")

# adapter codes of dropin_harness.cpp::dropin_optimize / ref_harness.cpp::ref_optimize
TNC, TNC_AD, TNC_AD_PARALLEL, FULL = 0, 1, 2, 100


def NLOPT(k, ad=False, parallel=False):
    return (30 if parallel else 20 if ad else 10) + k


def build():
    """(Re)build the drop-in artefacts where the reference sources exist; otherwise rely on the prebuilt ones."""
    if os.path.isdir("/root/reference/src"):
        from treepl_b200 import build as tbuild
        tbuild.build()
        root = os.path.dirname(HERE)
        subprocess.check_call(["make", "-s", "-C", os.path.join(root, "oracle"), "ref"])
        subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "dropin")])
    return all(os.path.exists(os.path.join(BUILD, f)) for f in ("libdropin_b200.so", "libdropin_ref.so", "treePL_b200"))


_libs = {}


def lib(kind):
    """kind: 'b200' (GPU engine) or 'ref' (the reference's CPU class behind the same C ABI)."""
    if kind not in _libs:
        L = C.CDLL(os.path.join(BUILD, "libdropin_%s.so" % kind))
        ip, dp, ci, cd = C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int, C.c_double
        L.dropin_optimize.restype = ci
        L.dropin_optimize.argtypes = [ci, ip, ip, ip, ip] + [dp] * 9 + [ci, cd, ci, ip, ci, ci, cd, cd, ci, ci, dp, ip, dp, dp, ci, ip, dp, dp,
                                                                    C.c_char_p, ci]
        _libs[kind] = L
    return _libs[kind]


def optimize(kind, flat, which, lf=False, smoothing=1.0, log_pen=False, cvnodes=None, numiter=10000, ftol=1e-7, xtol=1e-7,
             seed=1, x0=None, trace_cap=200000):
    """Run the reference adapter `which` through the C++ mirror on backend `kind`.
    -> dict(rc, x, f, trace, rates, dates, log, numparams)"""
    L = lib(kind)
    n = len(flat["parents"])
    ints = [np.ascontiguousarray(flat[k], dtype=np.int32) for k in FLAT_ORDER_I]
    dbls = [np.ascontiguousarray(flat[k], dtype=np.float64) for k in FLAT_ORDER_D]
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    cv = np.ascontiguousarray(cvnodes, dtype=np.int32) if cvnodes is not None else None
    x = np.zeros(3 * n + 2)
    if x0 is not None:
        x[: len(x0)] = x0
    nump, ntrace, f = C.c_int(), C.c_int(), C.c_double()
    trace = np.zeros(trace_cap)
    rates, dates = np.zeros(n), np.zeros(n)
    log = C.create_string_buffer(1 << 16)
    rc = L.dropin_optimize(n, *[ip(a) for a in ints], *[dp(a) for a in dbls], int(lf), float(smoothing), int(log_pen),
                           ip(cv) if cv is not None else None, int(which), int(numiter), float(ftol), float(xtol), int(seed),
                           int(x0 is not None), dp(x), C.byref(nump), C.byref(f), dp(trace), trace_cap, C.byref(ntrace),
                           dp(rates), dp(dates), log, len(log))
    text = log.value.decode(errors="replace")
    if rc == -1000:
        raise RuntimeError(text)
    return {"rc": rc, "x": x[: nump.value].copy(), "f": f.value, "trace": trace[: min(ntrace.value, trace_cap)].copy(),
            "ncalls": ntrace.value, "rates": rates, "dates": dates, "log": text, "numparams": nump.value}
