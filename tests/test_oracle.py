"""Pins the CPU oracle (oracle/pl_oracle.c) and the O(N) flattener against the reference.

Three anchors (the reference ships no tests of its own):
  1. the known-answer values of SURVEY.md §8(c), produced by a reference build;
  2. tests/golden/*.npz, produced by oracle/make_golden.py from the reference's own
     objects on the reference's own example trees;
  3. the live reference library (oracle/_ref/libtreepl_ref.so) when it is present.
"""
import math
import os

import numpy as np
import pytest

from golden_util import Golden, golden_names, rel_err
from oracle import plo, refbind
from treepl_b200 import treeflat as tf

TEST_TRE = "(((a:0.3,b:0.3):0.2,((c:0.2,d:0.2):0.1,e:0.3):0.2):0.1,(f:0.4,g:0.4):0.2):0.1;"
TEST_CALS = [(["a", "f"], 10, 10), (["a", "b"], 3, 4)]
KAT_DATES = [3.5426307087034665, 6.7707376136974116, 3.5239595263413004, 1.5246707634692467, 3.7124653890791839]
KAT_RATE = 0.078792967378301368


def kat_x():
    return np.array([KAT_RATE * (1 + 0.1 * math.sin(i)) for i in range(1, 13)] + KAT_DATES)


def kat_problem(log_pen=False):
    ft = tf.flatten_newick(TEST_TRE, 10000, TEST_CALS, seed=1)
    d = ft.as_dict()
    # the KAT was taken at the reference's own start state
    d["start_dates"] = np.array([10, KAT_DATES[0], 0, 0, KAT_DATES[1], KAT_DATES[2], 0, KAT_DATES[3], 0, 0,
                                 KAT_DATES[4], 0, 0], dtype=np.float64)
    o = plo.OracleProblem(d, log_pen=log_pen)
    o.set_smoothing(0.1)
    return ft, o


def test_flatten_test_tre_matches_survey_vectors():
    """index and constant contract on examples/test.tre."""
    ft = tf.flatten_newick(TEST_TRE, 10000, TEST_CALS, seed=1)
    assert ft.parents.tolist() == [-1, 0, 1, 1, 0, 4, 5, 5, 7, 7, 4, 10, 10]
    assert ft.free.tolist() == [0, 1, 0, 0, 1, 1, 0, 1, 0, 0, 1, 0, 0]
    assert ft.c.tolist() == [1000, 2000, 4000, 4000, 1000, 2000, 3000, 1000, 2000, 2000, 2000, 3000, 3000]
    off = ft.child_offsets()
    kids = [ft.children[off[i]:off[i + 1]].tolist() for i in range(13)]
    assert kids == [[4, 1], [3, 2], [], [], [10, 5], [7, 6], [], [9, 8], [], [], [12, 11], [], []]
    assert ft.min.tolist() == [10, 0, 0, 0, 3, 0, 0, 0, 0, 0, 3, 0, 0]
    assert ft.max.tolist() == [10, 10, 1e20, 1e20, 10, 10, 1e20, 10, 1e20, 1e20, 4, 1e20, 1e20]
    assert ft.pen_min[10] == 3 and ft.pen_max[10] == 4
    assert (np.delete(ft.pen_min, 10) == -1).all() and (np.delete(ft.pen_max, 10) == -1).all()
    assert [ft.names[i] for i in (2, 3, 6, 8, 9, 11, 12)] == list("gfedcba")
    assert tf.log_fact(1.0) == -0.081061889127438169
    fp, P = tf.generate_param_order_vector(ft.free, lf=False)
    assert P == 17
    assert fp.tolist() == [-1] + list(range(12)) + [-1, 12, -1, -1, 13, 14, -1, 15, -1, -1, 16, -1, -1]


def test_oracle_known_answers_pl():
    """SURVEY.md §8(c) golden vector: f bit-exact, both gradients."""
    ft, o = kat_problem()
    o.set_freeparams(False)
    x = kat_x()
    f, g_an = o.calc_pl_and_gradient(x)
    assert f == 239399.94428679207
    f_ad, g_ad = o.calc_function_gradient(x)
    assert abs(f_ad - 239399.94428679207) <= 1e-10
    want = np.array([-23406.393408187661, -46531.032327605913, -50055.970599193817, -13727.39652108842,
                     -28071.929595106591, -39165.378850120374, -11907.075866837093, -23096.253066229467,
                     -24376.776861839076, -26840.249346265493, -42301.202220029205, -40229.55996358618,
                     -1948.4047532878931, -960.21815365522548, -735.40837417371245, -2123.254686605891,
                     -962.11637706838883])
    assert rel_err(g_ad, want) < 1e-14
    want_an = want.copy()
    want_an[-1] = -962.14169046341976           # the analytic gradient omits the barrier term
    assert rel_err(g_an, want_an) < 1e-14


def test_oracle_known_answers_cv_lf_logpen():
    ft, o = kat_problem()
    cv = np.zeros(13, np.int32)
    cv[12] = 1
    o.set_cv(cv)
    o.set_freeparams(False, use_cv=True)
    assert o.numparams == 16
    x = np.delete(kat_x(), 11)
    assert o.calc_pl(x) == 214522.48507340456
    # LF
    ft, o = kat_problem()
    o.set_freeparams(True)
    x = np.array([KAT_RATE] + KAT_DATES)
    f, g_an = o.calc_pl_and_gradient(x)
    assert f == 239340.0552474407
    assert abs(g_an[0] / -380706.56554890471 - 1) < 1e-15
    f_ad, g_ad = o.calc_function_gradient(x)
    assert abs(g_ad[0] / -368014.07754847454 - 1) < 1e-14
    # log_pen
    ft, o = kat_problem(log_pen=True)
    o.set_freeparams(False)
    f, g_an = o.calc_pl_and_gradient(kat_x())
    assert f == 239399.94491876027
    f_ad, g_ad = o.calc_function_gradient(kat_x())
    assert abs(g_ad[0] / -23406.300684481106 - 1) < 1e-14 and abs(g_ad[3] / -13727.505386851408 - 1) < 1e-14
    assert abs(g_an[0] / -23406.159931587172 - 1) < 1e-14
    o.set_cv(np.eye(13, dtype=np.int32)[12])
    o.set_freeparams(False, use_cv=True)
    assert o.calc_pl(np.delete(kat_x(), 11)) == 214522.48570537276


def test_oracle_barrier_reset_vs_skip():
    """a date exactly ON a penalised bound; calc_pl resets the whole
    barrier accumulator there, the AD paths only skip the term."""
    cals = TEST_CALS + [(["c", "d"], 1, 2)]
    ft = tf.flatten_newick(TEST_TRE, 10000, cals, seed=1)
    d = ft.as_dict()
    o = plo.OracleProblem(d)
    o.set_smoothing(0.1)
    x0 = o.set_freeparams(False)
    fp = o.ints("freeparams")
    x = kat_x()
    x[fp[13 + 7]] = 1.5       # MRCA(c,d) = node 7 in [1,2]
    x[fp[13 + 10]] = 3.0      # node 10 exactly on pen_min = 3
    f = o.calc_pl(x)
    f_ad, _ = o.calc_function_gradient(x)
    assert abs((f_ad - f) - 0.0025 * (1 / 0.5 + 1 / 0.5)) < 1e-9


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_goldens(name):
    """Every recorded evaluation of the reference: f bit-exact, analytic gradient bit-exact,
    AD gradient to 1e-12 (the tape's accumulation order is not reproduced)."""
    gd = Golden(name)
    probs = {}
    for e in gd.evals:
        key = e["log_pen"]
        if key not in probs:
            probs[key] = plo.OracleProblem(gd.flat, log_pen=bool(key))
        o = probs[key]
        o.set_state(gd.flat["start_rates"], gd.flat["start_dates"])
        o.set_smoothing(e["smoothing"])
        cv = gd.arr(e["k"], "cv")
        o.set_cv(cv)
        o.set_freeparams(bool(e["lf"]), use_cv=cv is not None)
        assert o.numparams == e["numparams"]
        assert np.array_equal(o.ints("freeparams"), gd.arr(e["k"], "fp"))
        x = gd.arr(e["k"], "x")
        f, g = o.calc_pl_and_gradient(x)
        want = float(e["f_pl"])
        assert f == want, (name, e, f)
        g_an = gd.arr(e["k"], "g_an")
        if g_an is not None:
            assert np.array_equal(g, g_an, equal_nan=True), (name, e["k"], e["label"])
            dur = gd.arr(e["k"], "durations")
            if dur is not None:
                assert np.array_equal(o.doubles("durations"), dur)
        if "f_ad" in e:
            f_ad, g_ad = o.calc_function_gradient(x)
            want_ad = float(e["f_ad"])
            if math.isfinite(want_ad):
                assert abs(f_ad - want_ad) <= 1e-13 * abs(want_ad), (name, e, f_ad)
            else:
                assert f_ad == want_ad or (math.isnan(f_ad) and math.isnan(want_ad))
            ref_g = gd.arr(e["k"], "g_ad")
            if ref_g is not None and math.isfinite(want_ad) and want_ad < 1e15:
                assert rel_err(g_ad, ref_g) < 1e-12, (name, e["k"], e["label"], rel_err(g_ad, ref_g))


@pytest.mark.parametrize("name", golden_names())
def test_flattener_matches_reference_goldens(name):
    """O(N) flattener == the reference's extract_tree_info on the reference's example trees."""
    gd = Golden(name)
    cals = [(t, lo, hi) for t, lo, hi in gd.meta["calibrations"]]
    ft = tf.flatten_newick(gd.meta["newick"], gd.meta["numsites"], cals, seed=1)
    for k in ("parents", "child_counts", "children", "free"):
        assert np.array_equal(getattr(ft, k), gd.flat[k]), k
    for k in ("c", "lf", "min", "max", "pen_min", "pen_max"):
        assert np.array_equal(getattr(ft, k), gd.flat[k], equal_nan=True), k
    assert ft.names == gd.meta["names"]
    # start point is feasible under the oracle (start points are not part of the parity contract)
    o = plo.OracleProblem(ft.as_dict())
    x0 = o.set_freeparams(True)
    assert o.calc_pl(x0) < 1e15


@pytest.mark.skipif(not refbind.available(), reason="oracle/_ref not built (make -C oracle ref)")
def test_oracle_matches_live_reference_random_points():
    rng = np.random.RandomState(7)
    gd = Golden("M1155")
    ref = refbind.RefProblem.from_flat(gd.flat)
    o = plo.OracleProblem(gd.flat)
    n = o.numnodes
    tips = np.flatnonzero(gd.flat["child_counts"] == 0)
    for trial in range(20):
        lam = 10.0 ** rng.uniform(-3, 3)
        cv = None
        if trial % 2:
            cv = np.zeros(n, np.int32)
            cv[rng.choice(tips, 1 + trial % 5, replace=False)] = 1
        for q in (ref, o):
            q.set_state(gd.flat["start_rates"], gd.flat["start_dates"])
            q.set_smoothing(lam)
            q.set_cv(cv)
        x0 = ref.set_freeparams(False, cv is not None)
        x1 = o.set_freeparams(False, cv is not None)
        assert np.array_equal(x0, x1)
        nr = n - 1 - (int(cv.sum()) if cv is not None else 0)
        x = x0.copy()
        x[:nr] = 0.01 * np.exp(rng.randn(nr))
        x[nr:] *= 1 - 1e-3 * rng.rand(x.size - nr)
        f_r, g_r = ref.calc_pl_and_gradient(x)
        f_o, g_o = o.calc_pl_and_gradient(x)
        assert f_r == f_o
        if f_r < 1e15:
            assert np.array_equal(g_r, g_o)
            fa_r, ga_r = ref.calc_function_gradient(x)
            fa_o, ga_o = o.calc_function_gradient(x)
            assert abs(fa_r - fa_o) <= 1e-13 * abs(fa_r)
            assert rel_err(ga_o, ga_r) < 1e-12


def test_synthetic_problem_is_feasible_and_consistent():
    ft, nw, cals = tf.synth_problem(300, seed=3, ncal=4)
    assert ft.numnodes == 599
    assert (ft.pen_min != -1).sum() >= 1
    o = plo.OracleProblem(ft.as_dict())
    o.set_smoothing(1.0)
    x0 = o.set_freeparams(False)
    fp, P = tf.generate_param_order_vector(ft.free, lf=False)
    assert P == o.numparams and np.array_equal(fp, o.ints("freeparams"))
    x0[: ft.numnodes - 1] = ft.start_rates[1] * ft.numsites
    f = o.calc_pl(x0)
    assert np.isfinite(f) and f < 1e15
