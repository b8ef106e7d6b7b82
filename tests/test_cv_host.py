"""Host logic of the CV driver (treepl_b200/cv.py) on the CPU: fold sampling and chi-square scoring
(reference src/main.cpp:603-656, 737-793)."""
import numpy as np

from golden_util import Golden
from treepl_b200 import cv
from treepl_b200 import treeflat as tf


def _m1155():
    gd = Golden("M1155")
    cals = [(t, lo, hi) for t, lo, hi in gd.meta["calibrations"]]
    ft = tf.flatten_newick(gd.meta["newick"], gd.meta["numsites"], cals, seed=1)
    return gd.flat, [int(t) for t in ft.tips_postorder]


def test_random_groups_are_sister_free_and_avoid_root_children():
    flat, tips = _m1155()
    par = flat["parents"]
    groups = cv.sample_random_groups(flat, tips, n_groups=10, frac=0.1, seed=3)
    assert len(groups) == 10
    size = int(len(tips) * 0.1 + 0.5)
    for g in groups:
        assert len(g) == size and len(set(g)) == size
        assert all(par[t] != 0 for t in g)                       # main.cpp:621
        assert len({par[t] for t in g}) == size                  # no two sisters
        assert all(flat["child_counts"][t] == 0 for t in g)
    assert groups == cv.sample_random_groups(flat, tips, 10, 0.1, seed=3)      # reproducible


def test_chisq_terms_follow_the_reference_formula():
    """(c - r_parent d)^2 / (r_parent d); a child of the root is predicted with its sibling's rate (main.cpp:739-748);
    random-subsample folds average over the group (main.cpp:792)."""
    flat, tips = _m1155()
    n = len(flat["parents"])
    par = flat["parents"]
    fp, P = tf.generate_param_order_vector(flat["free"], lf=False)
    rng = np.random.RandomState(0)
    x = cv.start_vector(flat, P, n - 1)
    x[: n - 1] *= np.exp(0.2 * rng.randn(n - 1))
    dates = flat["start_dates"].copy()
    for i in range(n):
        if fp[n + i] != -1:
            dates[i] = x[fp[n + i]]
    want = []
    for t in tips[:12]:
        p = par[t]
        if p == 0:
            sib = [j for j in range(n) if par[j] == 0 and j != t][-1]
            r = x[fp[sib]]
        else:
            r = x[fp[p]]
        e = r * (dates[p] - dates[t])
        want.append((flat["c"][t] - e) ** 2 / e)
    got = [cv.group_chisq(flat, x, [t], fp, False) for t in tips[:12]]
    assert np.allclose(got, want, rtol=1e-14)
    grp = tips[:12]
    assert np.isclose(cv.group_chisq(flat, x, grp, fp, True), np.sum(want) / 12, rtol=1e-13)
    assert np.isclose(cv.group_chisq(flat, x, grp, fp, False), np.sum(want), rtol=1e-13)


def test_single_tip_scoring_of_all_lanes_equals_the_per_fold_formula():
    """cv.single_tip_chisq (LOOCV, every lane at once) == cv.group_chisq lane by lane, bit for bit, incl. a tip that is a
    child of the root."""
    flat, tips = _m1155()
    n = len(flat["parents"])
    fp, P = tf.generate_param_order_vector(flat["free"], lf=False)
    rng = np.random.RandomState(1)
    x = cv.start_vector(flat, P, n - 1)
    lane_tips = [int(t) for t in tips] + [int(t) for t in tips[:7]]
    X = x[:, None] * np.exp(0.1 * rng.randn(P, len(lane_tips)))
    got = cv.single_tip_chisq(flat, X, lane_tips, fp)
    want = np.array([cv.group_chisq(flat, X[:, b], [t], fp, False) for b, t in enumerate(lane_tips)])
    assert np.array_equal(got, want)


def test_probe_based_scoring_equals_the_per_fold_formula():
    """What the CV driver does since the fold optima stay on the device: cv.fold_probe_rows picks the X rows a fold's chi-square
    needs, cv.chisq_from_probes scores from them == cv.group_chisq on the whole vector, bit for bit (ragged random groups and
    single-tip folds, incl. children of the root)."""
    flat, tips = _m1155()
    n = len(flat["parents"])
    fp, P = tf.generate_param_order_vector(flat["free"], lf=False)
    rng = np.random.RandomState(2)
    x = cv.start_vector(flat, P, n - 1)
    rnd = cv.sample_random_groups(flat, tips, n_groups=4, frac=0.2, seed=5)
    for groups, randomcv in (([[int(t)] for t in tips], False), (rnd + [rnd[0][:3]], True)):
        B = len(groups)
        X = x[:, None] * np.exp(0.1 * rng.randn(P, B))
        rows, ptips = cv.fold_probe_rows(flat, groups, fp)
        probes = np.where(rows >= 0, X[np.maximum(rows, 0), np.arange(B)[None, :]], 0.0)       # what cols_probe_kernel returns
        got = cv.chisq_from_probes(flat, probes, rows, ptips, randomcv)
        want = np.array([cv.group_chisq(flat, X[:, b], groups[b], fp, randomcv) for b in range(B)])
        assert np.array_equal(got, want)


def test_smoothing_grid_follows_the_reference_fixups():
    """main.cpp:275-284 (swap start/stop, invert a step > 1 in float) and :668,814 (while(curcv >= cvstop) curcv *= step)."""
    import pytest
    from treepl_b200 import config
    g = config.smoothing_grid(1000.0, 0.1, 0.1)
    cur, want = 1000.0, []
    while cur >= 0.1:
        want.append(cur)
        cur *= 0.1
    assert g == want and len(g) in (4, 5)
    assert config.smoothing_grid(0.1, 1000.0, 0.1) == want                       # swapped
    g10 = config.smoothing_grid(1000.0, 0.1, 10)                                 # inverted in float: terminates
    assert 4 <= len(g10) <= 5 and abs(g10[1] - 100.0) < 1e-4 and g10[1] != 100.0
    assert config.Config(cvstart=0.0001, cvstop=100.0, cvmultstep=0.1).smoothing_grid()[0] == 100.0   # examples/pl4203.cppr8s
    for bad in (1.0, 0.0, -2.0):
        with pytest.raises(ValueError):
            config.smoothing_grid(1000.0, 0.1, bad)


def test_control_file_tokens_like_the_reference():
    """mrca / min / max values tokenise on commas and blanks (main.cpp:109-137); randomcv implies cv (:237-239);
    atoi / atof tolerate trailing text."""
    from treepl_b200 import config
    cfg = config.parse_config_text("treefile = t.tre\nnumsites = 2744 sites\nmrca = A t1,t2  t3\nmin = A,10\nmax = A 20.5x\n"
                                   "randomcv\n[cv]\n#thorough\ncvmultstep = 10\nsmooth = 1e-4\n")
    assert cfg.mrcas == {"A": ["t1", "t2", "t3"]} and cfg.mins == {"A": 10.0} and cfg.maxs == {"A": 20.5}
    assert cfg.randomcv and cfg.cv and not cfg.thorough
    assert cfg.numsites == 2744 and cfg.smooth == 1e-4
    assert len(cfg.smoothing_grid()) in (4, 5)
