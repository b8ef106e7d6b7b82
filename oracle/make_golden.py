"""TEST INFRASTRUCTURE ONLY: generate tests/golden/*.npz from the REFERENCE ITSELF.

Runs the unmodified reference objects (oracle/_ref/libtreepl_ref.so, built by
`make -C oracle ref` from /root/reference/src) on the reference's own example inputs
(/root/reference/examples) and records, for a set of seeded parameter vectors per
tree and per mode (LF / PL / CV-masked / log_pen; feasible, infeasible and edge points):

    f_pl, g_an   calc_pl + calc_gradient            (pl_calc_parallel.cpp:73, :657)
    f_ad, g_ad   calc_function_gradient (RAD tape)  (pl_calc_parallel.cpp:132)
    f_adolc, g_adolc   calc_pl_function_gradient_adolc / calc_lf_function_gradient_adolc (ADOL-C 2.6.3 tape, :622 / :224)

together with the flat arrays produced by the reference's extract_tree_info.
/root/reference does not exist on the GPU box, so these fixtures are what travels.

    python -m oracle.make_golden          # rewrites tests/golden/
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import refbind as rb  # noqa: E402
from treepl_b200 import config as cf  # noqa: E402

EX = "/root/reference/examples"
OUT = os.path.join(ROOT, "tests", "golden")

FLAT_KEYS = ("parents", "child_counts", "children", "free", "c", "lf", "min", "max", "pen_min", "pen_max",
             "start_dates", "start_rates", "start_durations")


def points(p, fl, lf, cv, rng, numsites, want_edge):
    """Seeded parameter vectors: a generic feasible point, a second one, then edge cases."""
    n = p.numnodes
    p.set_state(fl["start_rates"], fl["start_dates"])      # earlier evaluations mutated the member state
    x0 = p.set_freeparams(lf, cv is not None)
    P = p.numparams
    nr = 1 if lf else (n - 1 - (int(np.sum(cv)) if cv is not None else 0))
    base = fl["start_rates"][1] * numsites
    out = []
    x = x0.copy(); x[:nr] = base * (1 + 0.3 * rng.rand(nr)); out.append(("generic", x))
    x = x0.copy(); x[:nr] = base * np.exp(0.5 * rng.randn(nr)); out.append(("lognormal", x))
    if want_edge and P > nr:
        fp = p.ints("freeparams")
        par = fl["parents"]
        free_nodes = [i for i in range(n) if fp[n + i] != -1]
        # a free date pushed above its parent's date -> negative duration -> LARGE
        for i in free_nodes:
            if i != 0:
                x = out[0][1].copy(); pd = x[fp[n + par[i]]] if fp[n + par[i]] != -1 else fl["start_dates"][par[i]]
                x[fp[n + i]] = pd * 1.5 + 1.0
                out.append(("neg_duration_or_bound", x)); break
        # a free date EQUAL to its parent's date -> duration 0 -> DBL_MIN path
        for i in free_nodes:
            if i != 0 and fl["max"][i] >= (out[0][1][fp[n + par[i]]] if fp[n + par[i]] != -1 else fl["start_dates"][par[i]]):
                x = out[0][1].copy(); pd = x[fp[n + par[i]]] if fp[n + par[i]] != -1 else fl["start_dates"][par[i]]
                x[fp[n + i]] = pd
                # children of i must stay younger: only use if all children dates < pd (true: child < i < parent)
                out.append(("zero_duration", x)); break
        x = out[0][1].copy(); x[P - 1] = -1.0; out.append(("negative_param", x))
        x = out[0][1].copy(); x[0] = 0.0; out.append(("zero_rate", x))
        # date exactly ON a penalised bound (calc_pl resets the barrier accumulator, AD skips the term)
        for i in free_nodes:
            if fl["pen_min"][i] != -1:
                x = out[0][1].copy(); x[fp[n + i]] = fl["pen_min"][i]; out.append(("on_pen_min", x)); break
        for i in free_nodes:
            if fl["pen_max"][i] != -1:
                x = out[0][1].copy(); x[fp[n + i]] = fl["pen_max"][i]; out.append(("on_pen_max", x)); break
    return out


def run_case(name, newick, numsites, cals, smooth, rng, rad_ok=True, want_edge=True, small=False):
    arrays, meta = {}, {"name": name, "numsites": numsites, "evals": [],
                        "calibrations": [[list(t), lo, hi] for t, lo, hi in cals], "newick": newick}
    k = 0
    for log_pen in (False, True):
        p = rb.RefProblem.from_newick(newick, numsites, cals, seed=1, log_pen=log_pen)
        fl = p.flat()
        if not log_pen:
            for key in FLAT_KEYS:
                arrays["flat_" + key] = fl[key]
            meta["names"] = [p.name(i) for i in range(p.numnodes)]
        n = p.numnodes
        tips = np.flatnonzero(fl["child_counts"] == 0)
        for lf, usecv in ((True, False), (False, False), (False, True)):
            cv = None
            if usecv:
                cv = np.zeros(n, np.int32)
                cv[rng.choice(tips, max(1, len(tips) // 10), replace=False)] = 1
            p.set_smoothing(smooth)
            p.set_cv(cv)
            pts = points(p, fl, lf, cv, rng, numsites, want_edge and not log_pen)
            if small:       # large tree: keep the fixture small (one point per mode, LF only once)
                pts = pts[:1]
                if lf and log_pen:
                    continue
            for label, x in pts:
                f_pl, g_an = p.calc_pl_and_gradient(x)
                rec = {"k": k, "label": label, "lf": int(lf), "log_pen": int(log_pen), "cv": int(usecv),
                       "smoothing": smooth, "numparams": int(p.numparams), "f_pl": repr(float(f_pl))}
                arrays["e%d_x" % k] = x
                arrays["e%d_fp" % k] = p.ints("freeparams")
                if cv is not None:
                    arrays["e%d_cv" % k] = cv
                if f_pl < 1e15:
                    arrays["e%d_g_an" % k] = g_an
                    if not small:
                        arrays["e%d_durations" % k] = p.doubles("durations")
                if rad_ok and label != "negative_param":
                    f_ad, g_ad = p.calc_function_gradient(x)
                    rec["f_ad"] = repr(float(f_ad))
                    if not np.isnan(g_ad).all():
                        arrays["e%d_g_ad" % k] = g_ad
                    # the ADOL-C twin: recorded only where it differs from the RAD tape
                    f_dc, g_dc = p.calc_function_gradient_adolc(x)
                    rec["f_adolc"] = repr(float(f_dc))
                    same = np.array_equal(g_dc, g_ad, equal_nan=True)
                    rec["g_adolc_equals_g_ad"] = bool(same)
                    if not same and not np.isnan(g_dc).all():
                        arrays["e%d_g_adolc" % k] = g_dc
                meta["evals"].append(rec)
                k += 1
        p.close()
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-12s N=%-6d evals=%-3d %6.1f KB" % (name, n, k, os.path.getsize(path) / 1024))


def main():
    if not rb.available():
        raise SystemExit("build the reference first: make -C oracle ref")
    rng = np.random.RandomState(20261019)
    for cfgname in ("test", "pl4203", "M1155", "vib", "clock", "big_test"):
        cfg = cf.parse_config(os.path.join(EX, cfgname + ".cppr8s"))
        nw = open(cfg.tree_path()).read().strip().splitlines()[0]
        run_case(cfgname, nw, cfg.numsites, cfg.calibrations(), cfg.smooth, rng,
                 rad_ok=True, want_edge=cfgname != "big_test", small=cfgname == "big_test")
    # two bounded calibrations so that the barrier reset lands mid-sum
    nw = open(os.path.join(EX, "test.tre")).read().strip()
    cals = [(["a", "f"], 10, 10), (["a", "b"], 3, 4), (["c", "d"], 1, 2)]
    run_case("test_two_cal", nw, 10000, cals, 0.1, rng)


if __name__ == "__main__":
    main()
