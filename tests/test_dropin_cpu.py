"""The drop-in boundary on the CPU.  The reference's own optimiser adapters
(src/optimize_tnc.cpp + src/tnc.c, src/optimize_nlopt.cpp + NLopt 2.4.2), compiled UNMODIFIED against the engine's C++
mirror (treepl_b200/host/pl_calc_b200.h through the one-line shim), must behave exactly like the same adapters on the
reference's own class when the C ABI underneath is served by the reference's arithmetic (tests/dropin/tpl_over_ref.cpp):
identical optimiser return code, number of callbacks, final vector and objective — bit for bit, because every number that
crosses the boundary is the reference's.  This pins the HOST logic of the mirror (public data members, lazy read-back of
rates / dates / durations, smoothing assignment, bounds construction through `rates.size()` / `min->at()`) without a GPU.
The GPU engine behind the same mirror is compared call by call in tests/test_dropin_gpu.py."""
import os
import re

import numpy as np
import pytest

import dropin_util as du
from golden_util import Golden
from oracle import refbind as rb

pytestmark = pytest.mark.skipif(not du.build(), reason="drop-in artefacts not built (needs /root/reference)")


def _ref_problem(gd, lf, smoothing, log_pen=False):
    """-> reference problem and main()'s start: in PL mode every rate is the Langley-Fitch start rate (the flat
    start_rates carry it only at node 1, main.cpp:477; the LF phase spreads it, pl_calc_parallel.cpp:82-89)."""
    p = rb.RefProblem.from_flat(gd.flat, log_pen=log_pen)
    p.set_smoothing(smoothing)
    x0 = p.set_freeparams(lf)
    if not lf:
        x0[: len(gd.flat["parents"]) - 1] = gd.flat["start_rates"][1]
    return p, x0


@pytest.mark.parametrize("name,lf,smoothing", [("test", True, 1.0), ("test", False, 0.1), ("pl4203", False, 100.0),
                                               ("M1155", True, 1.0), ("M1155", False, 10.0), ("vib", False, 1.0)])
@pytest.mark.parametrize("which", [du.TNC, du.TNC_AD, du.TNC_AD_PARALLEL, du.NLOPT(1), du.NLOPT(2, ad=True), du.NLOPT(3)])
def test_reference_adapters_through_the_mirror_equal_the_reference(name, lf, smoothing, which):
    gd = Golden(name)
    p, x0 = _ref_problem(gd, lf, smoothing)
    rc, x, f, log = p.optimize(which, x0, numiter=2000)
    p.close()
    got = du.optimize("ref", gd.flat, which, lf=lf, smoothing=smoothing, numiter=2000, x0=x0)
    assert got["numparams"] == x0.size
    assert got["rc"] == rc
    assert got["log"] == log                                   # "nfeval: N rc: ..." lines of the adapter
    assert np.array_equal(got["x"], x)
    assert got["f"] == f
    if which in (du.TNC, du.TNC_AD, du.TNC_AD_PARALLEL):
        m = re.search(r"nfeval: (\d+)", log)
        assert m and got["ncalls"] >= int(m.group(1))          # + the initial calc_pl of main()


def test_state_members_read_back_like_the_reference():
    """rates / dates after an optimisation (main.cpp:679-681) and log_pen + CV-masked layouts."""
    gd = Golden("vib")
    tips = np.flatnonzero(gd.flat["child_counts"] == 0)
    cv = np.zeros(len(gd.flat["parents"]), np.int32)
    cv[tips[[3, 17]]] = 1
    p = rb.RefProblem.from_flat(gd.flat, log_pen=True)
    p.set_smoothing(10.0)
    p.set_cv(cv)
    x0 = p.set_freeparams(False, use_cv=True)
    rc, x, f, log = p.optimize(du.TNC, x0, numiter=500)
    rates, dates = p.doubles("rates"), p.doubles("dates")
    p.close()
    got = du.optimize("ref", gd.flat, du.TNC, smoothing=10.0, log_pen=True, cvnodes=cv, numiter=500)
    assert got["rc"] == rc and np.array_equal(got["x"], x) and got["f"] == f
    assert np.array_equal(got["rates"], rates) and np.array_equal(got["dates"], dates)


def test_optimize_full_through_the_mirror_equals_the_reference():
    """utils.cpp:671-763 (annealer + TNC rounds, libc rand() seeded as main() does): same random stream, same numbers."""
    gd = Golden("test")
    p, x0 = _ref_problem(gd, False, 0.1)
    rc, x, f, log = p.optimize(du.FULL, x0, seed=1)
    p.close()
    got = du.optimize("ref", gd.flat, du.FULL, smoothing=0.1, seed=1, x0=x0)
    assert np.array_equal(got["x"], x) and got["f"] == f and got["log"] == log


def test_drop_in_builds_do_not_contain_the_reference_hot_path():
    """pl_calc_parallel.cpp / myradops.cpp must not be linked into the GPU drop-in artefacts."""
    for fn in ("libdropin_b200.so", "treePL_b200"):
        out = os.popen("nm -C %s 2>/dev/null" % os.path.join(du.BUILD, fn)).read()
        assert "pl_calc_parallel::calc_pl" not in out and "ADcontext" not in out, fn
        assert "tpl_calc_pl" in out, fn
