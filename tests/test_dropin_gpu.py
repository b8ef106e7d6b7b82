"""The drop-in boundary on the GPU.  The reference's own optimiser adapters
(src/optimize_tnc.cpp:33-231 + src/tnc.c, src/optimize_nlopt.cpp:22-341 + NLopt 2.4.2) and its own CLI (src/main.cpp),
compiled UNMODIFIED against the engine's C++ mirror (tests/dropin/Makefile: one -include flag), run on the B200 engine and
are compared with the same adapters on the reference's CPU arithmetic behind the same mirror (libdropin_ref.so, which
tests/test_dropin_cpu.py proves bit-identical to the reference's own class).  Same start, same adapter: the sequence of
objective values the optimiser sees must agree call by call while the two runs are in lockstep, and the optimum reached
must be the same."""
import json
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import dropin_util as du
from golden_util import GOLDEN_DIR, Golden

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not du.build(), reason="drop-in artefacts not built (needs /root/reference)")]

CASES = [("test", 0.1), ("pl4203", 100.0), ("M1155", 10.0)]


def _start(gd):
    """main()'s PL start: dates from get_feasible_start_dates, every rate = the Langley-Fitch start rate
    (main.cpp:477, pl_calc_parallel.cpp:82-89)."""
    from treepl_b200 import engine as eng
    fp, P = eng.generate_param_order_vector(gd.flat["free"], lf=False)
    n = len(gd.flat["parents"])
    x0 = np.zeros(P)
    x0[: n - 1] = gd.flat["start_rates"][1]
    for i in range(n):
        if fp[n + i] >= 0:
            x0[fp[n + i]] = gd.flat["start_dates"][i]
    return x0


def _lockstep(tr_gpu, tr_ref, tol):
    """length of the common prefix of the two objective traces at relative tolerance tol"""
    m = min(len(tr_gpu), len(tr_ref))
    k = 0
    while k < m and abs(tr_gpu[k] - tr_ref[k]) <= tol * max(abs(tr_ref[k]), 1.0):
        k += 1
    return k


@pytest.mark.parametrize("name,smoothing", CASES + [("vib", 10.0)])
@pytest.mark.parametrize("which", [du.TNC, du.TNC_AD, du.TNC_AD_PARALLEL])
def test_reference_tnc_adapters_on_the_gpu_engine(name, smoothing, which):
    """optimize_plcp_tnc with function_plcp (calc_pl + analytic calc_gradient, optimize_tnc.cpp:75-105), with
    function_plcp_ad (RAD calc_function_gradient + the 1e-40 clamp, :33-52) and optimize_plcp_tnc_ad_parallel
    (function_plcp_ad_parallel -> calc_pl_function_gradient_adolc, :54-73): every callback of the reference's TNC
    lands on the B200 engine."""
    gd = Golden(name)
    x0 = _start(gd)
    ref = du.optimize("ref", gd.flat, which, smoothing=smoothing, numiter=10000, x0=x0)
    got = du.optimize("b200", gd.flat, which, smoothing=smoothing, numiter=10000, x0=x0)
    assert got["numparams"] == ref["numparams"] == x0.size
    assert abs(got["trace"][0] - ref["trace"][0]) <= 1e-12 * abs(ref["trace"][0])
    # Call by call the optimiser sees the reference's objective (1e-10) for as long as the two runs are in lockstep.
    # TNC's finite-difference Hessian products and its `<` tests on nearly equal values amplify last-bit differences
    # of f and g (the reference itself is not reproducible across thread counts), so the lockstep prefix
    # is finite by nature; measured on a B200 it is 37 - 133 calls, i.e. the start, every Newton direction and every
    # line search up to the adapter's own termination test for most cases.
    k = _lockstep(got["trace"], ref["trace"], 1e-10)
    assert k >= min(30, len(ref["trace"]), len(got["trace"])), (k, len(ref["trace"]), got["trace"][:k + 1], ref["trace"][:k + 1])
    if k == len(ref["trace"]) == len(got["trace"]):           # whole run in lockstep: same nfeval, same return code
        assert got["rc"] == ref["rc"] and got["log"] == ref["log"]
        assert abs(got["f"] - ref["f"]) <= 1e-8 * abs(ref["f"])
        assert np.allclose(got["x"], ref["x"], rtol=1e-6, atol=0)


@pytest.mark.parametrize("name,smoothing", [("test", 0.1), ("pl4203", 100.0), ("vib", 10.0)])
@pytest.mark.parametrize("which", [du.NLOPT(1), du.NLOPT(2, ad=True), du.NLOPT(3, parallel=True)])
def test_reference_nlopt_adapters_on_the_gpu_engine(name, smoothing, which):
    """optimize_plcp_nlopt / _ad / _ad_parallel (optimize_nlopt.cpp:116-341): LBFGS, TNEWTON_PRECOND_RESTART, MMA."""
    gd = Golden(name)
    x0 = _start(gd)
    ref = du.optimize("ref", gd.flat, which, smoothing=smoothing, numiter=10000, x0=x0)
    got = du.optimize("b200", gd.flat, which, smoothing=smoothing, numiter=10000, x0=x0)
    assert abs(got["trace"][0] - ref["trace"][0]) <= 1e-12 * abs(ref["trace"][0])
    k = _lockstep(got["trace"], ref["trace"], 1e-10)
    assert k >= min(8, len(ref["trace"])), (k, got["trace"][:k + 1], ref["trace"][:k + 1])
    assert got["rc"] == ref["rc"]
    if which == du.NLOPT(2, ad=True):
        # the preconditioned truncated Newton of NLopt converges (nlopt result 1 / 3): same optimum
        assert abs(got["f"] - ref["f"]) <= 1e-8 * abs(ref["f"]), (got["f"], ref["f"])


def test_member_state_after_an_optimisation_is_the_engines():
    """rates / dates read back through the mirror's public members after optimize_plcp_tnc (main.cpp:679-681) with a
    CV mask and the log penalty."""
    gd = Golden("vib")
    n = len(gd.flat["parents"])
    tips = np.flatnonzero(gd.flat["child_counts"] == 0)
    cv = np.zeros(n, np.int32)
    cv[tips[[3, 17]]] = 1
    from treepl_b200 import engine as eng
    fp, P = eng.generate_param_order_vector(gd.flat["free"], lf=False, cvnodes=cv)
    x0 = np.zeros(P)
    for i in range(n):
        if fp[i] >= 0:
            x0[fp[i]] = gd.flat["start_rates"][1]
        if fp[n + i] >= 0:
            x0[fp[n + i]] = gd.flat["start_dates"][i]
    ref = du.optimize("ref", gd.flat, du.TNC, smoothing=10.0, log_pen=True, cvnodes=cv, numiter=500, x0=x0)
    got = du.optimize("b200", gd.flat, du.TNC, smoothing=10.0, log_pen=True, cvnodes=cv, numiter=500, x0=x0)
    assert np.isfinite(ref["trace"][0])
    assert abs(got["trace"][0] - ref["trace"][0]) <= 1e-12 * abs(ref["trace"][0])
    assert _lockstep(got["trace"], ref["trace"], 1e-10) >= min(30, len(ref["trace"]))
    # the read-back state is the engine's scatter of the final vector (calc_pl :79-99)
    for i in range(1, n):
        if fp[i] >= 0:
            assert got["rates"][i] == got["x"][fp[i]]
        if fp[n + i] >= 0:
            assert got["dates"][i] == got["x"][fp[n + i]]


def _write_config(work, gd, extra):
    tre = os.path.join(work, "in.tre")
    open(tre, "w").write(gd.meta["newick"] + "\n")
    lines = ["treefile = " + tre, "numsites = " + str(gd.meta["numsites"])]
    cals = gd.meta["calibrations"]
    if isinstance(cals, str):
        cals = json.loads(cals.replace("'", '"'))
    for k, (tips, lo, hi) in enumerate(cals):
        lines.append("mrca = N%d %s" % (k, " ".join(tips)))
        lines.append("min = N%d %r" % (k, lo))
        lines.append("max = N%d %r" % (k, hi))
    lines += extra + ["outfile = " + os.path.join(work, "out.tre"), "cvoutfile = " + os.path.join(work, "cv.out"), "seed = 1",
                      "nthreads = 1"]
    cfg = os.path.join(work, "run.cppr8s")
    open(cfg, "w").write("\n".join(lines) + "\n")
    return cfg


def _run_cli(gd, extra):
    with tempfile.TemporaryDirectory() as work:
        cfg = _write_config(work, gd, extra)
        res = subprocess.run([os.path.join(du.BUILD, "treePL_b200"), cfg], cwd=work, stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True, timeout=900)
        assert res.returncode == 0, res.stdout[-2000:]
        log = res.stdout
        lf = [float(v) for v in re.findall(r"lf calc: ([-+0-9.eE]+)", log)]
        pl = [float(v) for v in re.findall(r"after opt calc: ([-+0-9.eE]+)", log)]
        assert lf and pl, log[-2000:]
        tree = open(os.path.join(work, "out.tre")).read().strip()
    return lf[-1], pl[-1], tree


def test_reference_cli_runs_on_the_gpu_engine():
    """src/main.cpp itself (LF phase -> optimize_full with its annealer, TNC and NLopt rounds -> PL phase -> dated tree),
    linked against the B200 engine instead of pl_calc_parallel.cpp, on examples/test.cppr8s: the Langley-Fitch value,
    the final PL value and the dated tree are the reference binary's (tests/golden/e2e_reference.json)."""
    want = json.load(open(os.path.join(GOLDEN_DIR, "e2e_reference.json")))["test"]
    lf, pl, tree = _run_cli(Golden("test"), ["smooth = 0.1"])
    assert abs(lf - want["lf"]) <= 1e-6 * want["lf"], (lf, want["lf"])
    assert abs(pl - want["final_pl"]) <= 1e-4 * want["final_pl"], (pl, want["final_pl"])
    num = lambda s: np.array([float(v) for v in re.findall(r":([0-9.]+)", s)])
    assert re.sub(r":[0-9.]+", "", tree) == re.sub(r":[0-9.]+", "", want["dated_tree"])
    assert np.max(np.abs(num(tree) - num(want["dated_tree"]))) <= 2e-2, (tree, want["dated_tree"])


@pytest.mark.parametrize("name,extra,k", [
    ("pl4203", ["smooth = 100", "thorough", "lfiter = 1", "opt = 5", "optad = 2", "moredetailad"], 0),
    ("M1155", ["smooth = 1000", "opt = 2", "optad = 0", "moredetailad"], 0)])
def test_reference_cli_full_fit_equals_the_reference_run(name, extra, k):
    """The reference CLI on the GPU engine with the options of examples/<name>.cppr8s minus the CV loop, at the first
    smoothing value of that file's grid: LF optimum and PL optimum equal the reference binary's own run
    (`lf calc` and the first `after opt calc` line of its CV loop, which is this very fit — same start, same libc rand()
    stream position, main.cpp:497-575,668-678)."""
    want = json.load(open(os.path.join(GOLDEN_DIR, "e2e_reference.json")))[name]
    lf, pl, tree = _run_cli(Golden(name), extra)
    assert abs(lf - want["lf"]) <= 1e-6 * want["lf"], (lf, want["lf"])
    assert abs(pl - want["after_opt_calc"][k]) <= 1e-6 * want["after_opt_calc"][k], (pl, want["after_opt_calc"][k])
