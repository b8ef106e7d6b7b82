"""O(N) problem flattening: Newick -> the flat arrays the PL engine consumes.

This is the DATA CONTRACT of the hot path: every index the engine
sees is defined here and must equal the reference's bit for bit.  The reference builds
the same arrays in O(N^2) (Tree::getNodeByNodeNumber, tree.cpp:86, called per node from
utils.cpp:186,195,246,363); this module does it in O(N) with explicit stacks, so a
100k-tip tree flattens in seconds instead of ~11 minutes.

Reference behaviour mirrored (file:line under /root/reference/src):
  * Newick grammar of TreeReader::readTree            tree_reader.cpp:22-139
  * externalNodes / nodes = recursive post-order       tree.cpp:311-326
  * process_initial_branch_lengths (BL floor 1/numsites) utils.cpp:115-165 (collapse unsupported)
  * apply_node_label_preorder (explicit stack, children pushed 0..k-1, so the LAST child
    is numbered first)                                 utils.cpp:167-180
  * calc_char_durations / round / logFact              utils.cpp:182-190, 397-400, 226-228
  * calibration flags as main() sets them              main.cpp:336-387
  * set_mins_maxs                                      utils.cpp:27-66
  * setup_date_constraints                             utils.cpp:192-224
  * extract_tree_info                                  utils.cpp:230-292
  * generate_param_order_vector                        utils.cpp:295-340
  * get_feasible_start_dates / a_feasible_time / set_node_order / get_start_rate
                                                       utils.cpp:342-423, 98-109
    (same formulas; the random draws come from numpy instead of libc rand(), start
    points are not part of the parity contract - SURVEY.md §2 "Start-point generation")
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

DBL_MIN = 2.2250738585072014e-308


# --------------------------------------------------------------------------------------
# Newick
# --------------------------------------------------------------------------------------
_STOP = set(",):;[")


def parse_newick(text):
    """-> (parent, children, name, bl) in node-creation order; node 0 is the root.

    Follows TreeReader::readTree (tree_reader.cpp:22-139) token for token: '(' opens a
    child (or the root), ',' and ')' climb to the parent, a label after ')' names the
    internal node, ':' reads a branch length with strtod semantics, '[...]' is a comment,
    blanks are skipped and anything else starts a tip label.
    """
    s = text.strip()
    s = s.rstrip(" \n\r\t")
    if not s or s[-1] != ";":
        raise ValueError("Tree is invalid: missing concluding semicolon")
    parent, children, name, bl = [], [], [], []

    def new_node(p):
        parent.append(p)
        children.append([])
        name.append("")
        bl.append(0.0)
        k = len(parent) - 1
        if p >= 0:
            children[p].append(k)
        return k

    n = len(s)
    x = 0
    cur = -1
    started = False
    while True:
        ch = s[x]
        if ch == "(":
            if not started:
                cur = new_node(-1)
                started = True
            else:
                cur = new_node(cur)
        elif ch == ",":
            cur = parent[cur]
        elif ch == ")":
            cur = parent[cur]
            x += 1
            y = x
            while y < n and s[y] not in _STOP:
                y += 1
            name[cur] = s[x:y]
            x = y - 1
        elif ch == ";":
            break
        elif ch == ":":
            x += 1
            y = x + 1
            while y < n and s[y] not in _STOP:
                y += 1
            bl[cur] = _strtod(s[x:y])
            x = y - 1
        elif ch == "[":
            y = s.index("]", x + 1)
            x = y
        elif ch == " ":
            pass
        else:
            cur = new_node(cur)
            y = x + 1
            while y < n and s[y] not in ",):[":
                y += 1
            name[cur] = s[x:y]
            x = y - 1
        if x < n - 1:
            x += 1
    return parent, children, name, bl


def _strtod(tok):
    """strtod(): longest valid numeric prefix, 0.0 if none."""
    try:
        return float(tok)
    except ValueError:
        import re

        m = re.match(r"\s*[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)", tok)
        return float(m.group(0)) if m else 0.0


def _round_half_up(x):
    """the reference's own round() (utils.cpp:397-400): frac >= .5 -> ceil else floor."""
    t = x - int(x)
    return math.ceil(x) if t >= 0.5 else math.floor(x)


def log_fact(k):
    """Stirling with the reference's truncated 2*pi (utils.cpp:226-228)."""
    if k <= 0:
        return float("nan")
    return k * (math.log(k) - 1) + math.log(math.sqrt(6.28318 * k))


# --------------------------------------------------------------------------------------
# flat problem
# --------------------------------------------------------------------------------------
@dataclass
class FlatTree:
    """Flat arrays in REFERENCE node numbering (root = 0)."""

    parents: np.ndarray          # int32[N], parents[0] = -1
    child_counts: np.ndarray     # int32[N]
    children: np.ndarray         # int32[N-1]  children of node 0, node 1, ... in Newick order
    free: np.ndarray             # int32[N]
    c: np.ndarray                # float64[N]  char_durations
    lf: np.ndarray               # float64[N]  log_fact_char_durations
    min: np.ndarray              # float64[N]
    max: np.ndarray              # float64[N]
    pen_min: np.ndarray          # float64[N]  -1 = none
    pen_max: np.ndarray          # float64[N]  -1 = none
    start_dates: np.ndarray      # float64[N]
    start_rates: np.ndarray      # float64[N]
    start_durations: np.ndarray  # float64[N]
    names: list = field(default_factory=list)          # by node number
    tips_postorder: np.ndarray = None                  # node numbers of externalNodes (LOOCV fold order, main.cpp:603-608)
    bl: np.ndarray = None                              # floored branch lengths by node number
    numsites: int = 0

    @property
    def numnodes(self):
        return int(self.parents.shape[0])

    def as_dict(self):
        return {k: getattr(self, k) for k in ("parents", "child_counts", "children", "free", "c", "lf", "min", "max",
                                              "pen_min", "pen_max", "start_dates", "start_rates", "start_durations")}

    def child_offsets(self):
        off = np.zeros(self.numnodes + 1, dtype=np.int32)
        np.cumsum(self.child_counts, out=off[1:])
        return off


def generate_param_order_vector(free, lf, cvnodes=None):
    """utils.cpp:295-340 -> (freeparams int32[2N], numparams)."""
    free = np.asarray(free)
    n = free.shape[0]
    fp = np.full(2 * n, -1, dtype=np.int32)
    cv = cvnodes is not None and len(cvnodes) == n
    if lf and not cv:
        fp[1:n] = 0
        icount = 1
    elif cv:
        keep = np.asarray(cvnodes)[1:] == 0
        idx = np.cumsum(keep) - 1
        fp[1:n] = np.where(keep, idx, -1)
        icount = int(keep.sum())
        if lf:
            icount += 1
    else:
        fp[1:n] = np.arange(n - 1, dtype=np.int32)
        icount = n - 1
    isfree = free == 1
    fp[n:] = np.where(isfree, icount + np.cumsum(isfree) - 1, -1)
    return fp, icount + int(isfree.sum())


def initial_params(flat_or_rates, dates=None, freeparams=None, lf=False):
    """set_freeparams' emission of x0 (pl_calc_parallel.cpp:930-951)."""
    rates = flat_or_rates
    n = len(rates)
    fp = np.asarray(freeparams)
    out = []
    if lf:
        for i in range(n):
            if fp[i] != -1:
                out.append(rates[i])
                break
    else:
        out.extend(np.asarray(rates)[fp[:n] != -1].tolist())
    out.extend(np.asarray(dates)[fp[n:] != -1].tolist())
    return np.array(out, dtype=np.float64)


def flatten_newick(text, numsites, calibrations=(), seed=1, set1=False):
    """Newick + calibrations -> FlatTree (reference numbering).

    calibrations: iterable of (tip_names, min_or_None, max_or_None); the MRCA of the tips
    is the calibrated node (main.cpp:336-387).
    """
    parent, children, name, bl = parse_newick(text)
    n = len(parent)
    is_tip = [len(ch) == 0 for ch in children]

    # post-order (recursive order of tree.cpp:311-326), iterative
    post = []
    stack = [(0, 0)]
    while stack:
        v, k = stack.pop()
        if k < len(children[v]):
            stack.append((v, k + 1))
            stack.append((children[v][k], 0))
        else:
            post.append(v)

    # process_initial_branch_lengths (utils.cpp:115-137): floor every BL (root too) to 1/numsites
    if set1:
        bl = [1.0] * n
    else:
        floor = 1.0 / numsites
        bl = [floor if b < floor else b for b in bl]

    # calibrations (main.cpp:336-387): all mins first, then all maxs
    depth = [0] * n
    for v in reversed(post):      # reversed post-order visits parents before children
        if parent[v] >= 0:
            depth[v] = depth[parent[v]] + 1
    tip_by_name = {}
    for v in post:
        if is_tip[v]:
            tip_by_name[name[v]] = v  # last one wins, as Tree::getExternalNode(string) does (tree.cpp:48-55)

    def mrca(tips):
        nodes = [tip_by_name[t] for t in tips]
        a = nodes[0]
        for b in nodes[1:]:
            while a != b:
                if depth[a] >= depth[b]:
                    a = parent[a]
                else:
                    b = parent[b]
        return a

    mins, maxs = {}, {}
    minb = [False] * n
    maxb = [False] * n
    nmin = [0.0] * n
    nmax = [0.0] * n
    pen_minb = [False] * n
    pen_maxb = [False] * n
    pen_min = [0.0] * n
    pen_max = [0.0] * n
    for tips, lo, hi in calibrations:
        if lo is None:
            continue
        m = mrca(tips)
        mins[m] = float(lo); nmin[m] = float(lo); minb[m] = True; pen_minb[m] = True; pen_min[m] = float(lo)
    for tips, lo, hi in calibrations:
        if hi is None:
            continue
        m = mrca(tips)
        maxs[m] = float(hi); nmax[m] = float(hi); maxb[m] = True
        if minb[m]:
            if nmin[m] != nmax[m]:
                pen_maxb[m] = True; pen_max[m] = float(hi)
            else:
                pen_minb[m] = False; pen_max[m] = 0.0
        else:
            pen_maxb[m] = True; pen_max[m] = float(hi)

    # set_mins_maxs (utils.cpp:27-66), nodes in post-order.  `explicit_max` = calibrated
    # ancestors only: ancestors come later in post-order, so only user maxes are visible.
    explicit_max = dict(maxs)
    anc_max = [10000.0] * n   # min over explicit maxes of strict ancestors (default LARGE = 10000)
    for v in reversed(post):
        p = parent[v]
        if p >= 0:
            a = anc_max[p]
            if p in explicit_max and explicit_max[p] < a:
                a = explicit_max[p]
            anc_max[v] = a
    for v in post:
        if is_tip[v]:
            continue
        if v not in mins:
            ymin = 0.0
            for ch in children[v]:
                if not is_tip[ch]:
                    cm = mins.setdefault(ch, 0.0)   # operator[] inserts 0 (utils.cpp:37)
                    if cm > ymin:
                        ymin = cm
            mins[v] = ymin; nmin[v] = ymin; minb[v] = True
        if v not in maxs:
            maxs[v] = anc_max[v]; nmax[v] = anc_max[v]; maxb[v] = True

    # setup_date_constraints (utils.cpp:192-224)
    free = [False] * n
    date = [0.0] * n
    for v in range(n):
        if is_tip[v]:
            date[v] = 0.0; free[v] = False; minb[v] = False; maxb[v] = False
        elif v in mins or v in maxs:
            free[v] = True
            if v in mins:
                nmin[v] = mins[v]
            if v in maxs:
                nmax[v] = maxs[v]
            if v in mins and v in maxs and mins[v] == maxs[v]:
                date[v] = mins[v]; free[v] = False
        else:
            free[v] = True

    # apply_node_label_preorder (utils.cpp:167-180)
    number = [0] * n
    st = [0]
    cnt = 0
    order = []
    while st:
        v = st.pop()
        number[v] = cnt
        order.append(v)
        cnt += 1
        st.extend(children[v])
    # `order[k]` = creation index of the node numbered k

    N = n
    f_par = np.empty(N, dtype=np.int32)
    f_cc = np.empty(N, dtype=np.int32)
    f_free = np.empty(N, dtype=np.int32)
    f_c = np.empty(N); f_lf = np.empty(N); f_min = np.empty(N); f_max = np.empty(N)
    f_pmin = np.empty(N); f_pmax = np.empty(N); f_bl = np.empty(N)
    f_children = []
    names = [""] * N
    for k, v in enumerate(order):
        f_par[k] = -1 if parent[v] < 0 else number[parent[v]]
        f_cc[k] = len(children[v])
        f_children.extend(number[ch] for ch in children[v])
        f_free[k] = 1 if free[v] else 0
        cd = float(_round_half_up(bl[v] * numsites))
        f_c[k] = cd
        f_lf[k] = log_fact(cd)
        f_min[k] = nmin[v] if minb[v] else 0.0
        f_max[k] = nmax[v] if maxb[v] else 1e20
        f_pmax[k] = pen_max[v] if pen_maxb[v] else -1.0
        f_pmin[k] = pen_min[v] if pen_minb[v] else -1.0
        f_bl[k] = bl[v]
        names[k] = name[v]
    tips_post = np.array([number[v] for v in post if is_tip[v]], dtype=np.int32)

    ft = FlatTree(parents=f_par, child_counts=f_cc, children=np.array(f_children, dtype=np.int32), free=f_free,
                  c=f_c, lf=f_lf, min=f_min, max=f_max, pen_min=f_pmin, pen_max=f_pmax,
                  start_dates=np.zeros(N), start_rates=np.zeros(N), start_durations=np.zeros(N),
                  names=names, tips_postorder=tips_post, bl=f_bl, numsites=numsites)
    fixed_dates = np.array([date[v] for v in order])
    feasible_start(ft, fixed_dates, seed=seed)
    return ft


def feasible_start(ft: FlatTree, fixed_dates, seed=1):
    """get_feasible_start_dates + a_feasible_time + set_node_order + get_start_rate
    (utils.cpp:342-423, 98-109), O(N), numpy RNG instead of libc rand()."""
    rng = np.random.RandomState(seed)
    N = ft.numnodes
    par = ft.parents
    off = ft.child_offsets()
    ch = ft.children
    # node order = max #edges down to a tip (utils.cpp:410-423); children have larger
    # preorder numbers than their parent, so one reverse sweep suffices
    order = np.zeros(N, dtype=np.int64)
    for i in range(N - 1, 0, -1):
        p = par[i]
        if order[p] < order[i] + 1:
            order[p] = order[i] + 1
    dates = np.array(fixed_dates, dtype=np.float64).copy()
    u = rng.random_sample(N)
    if ft.free[0] == 1:
        lo, hi = ft.min[0], ft.max[0]
        dates[0] = lo + (hi - lo) * (0.02 + u[0] * 0.96) if hi < 10000 else lo * 1.25
    # a_feasible_time recursion in child order == any parent-before-child order for the values
    for i in range(1, N):
        if ft.free[i] == 1:
            t_anc = dates[par[i]]
            if ft.max[i] < t_anc:      # maxb is true for every internal node (utils.cpp:57-58)
                t_anc = ft.max[i]
            dates[i] = t_anc - (t_anc - ft.min[i]) * (0.02 + u[i] * 0.96) / math.log(order[i] + 3)
        elif ft.child_counts[i] > 0:
            pass                        # fixed-age internal node keeps its date
    dur = np.empty(N)
    dur[0] = -1.0
    dur[1:] = dates[par[1:]] - dates[1:]
    dur[dur == 0] = DBL_MIN
    ft.start_dates = dates
    ft.start_durations = dur
    ft.start_rates = np.zeros(N)
    ft.start_rates[1] = float(np.sum(ft.bl)) / float(np.sum(dur))   # get_start_rate; written to slot 1 only (main.cpp:477)
    return ft


# --------------------------------------------------------------------------------------
# synthetic trees
# --------------------------------------------------------------------------------------
def synth_yule_newick(ntips, seed=1, root_age=100.0, rate=0.002, sigma=0.3, floor=1e-6):
    """Pure-birth (Yule) tree with `ntips` tips, ultrametric in time, scaled so the root
    age is `root_age`; branch lengths = time x rate x lognormal(0, sigma) substitutions per
    site, floored.  Returns (newick, node_ages dict by tip-set free representation is not
    needed: ages are re-derivable).  O(N)."""
    rng = np.random.RandomState(seed)
    # grow: active lineages list; pick random lineage to split, waiting time Exp(k)
    left = [-1, -1]; right = [-1, -1]
    birth = [0.0]      # time (from root, forward) at which node was born ; node 0 = root
    parent = [-1]
    left = [-1]; right = [-1]
    active = [0]
    t = 0.0
    split_time = [0.0]
    while len(active) < ntips:
        k = len(active)
        t += rng.exponential(1.0 / k)
        j = rng.randint(k)
        v = active[j]
        split_time[v] = t
        a = len(parent); b = a + 1
        parent.extend([v, v]); left.extend([-1, -1]); right.extend([-1, -1]); split_time.extend([0.0, 0.0])
        left[v] = a; right[v] = b
        active[j] = a
        active.append(b)
    t += rng.exponential(1.0 / len(active))
    total = t - split_time[0]
    n = len(parent)
    age = [0.0] * n            # time before present, scaled
    for v in range(n):
        if left[v] >= 0:
            age[v] = (t - split_time[v]) / total * root_age
    eps = rng.lognormal(0.0, sigma, size=n)
    bl = [0.0] * n
    for v in range(1, n):
        bl[v] = max(floor, (age[parent[v]] - age[v]) * rate * eps[v])
    # iterative Newick writer
    out = []
    stack = [(0, 0)]
    tipno = 0
    while stack:
        v, st = stack.pop()
        if left[v] < 0:
            out.append("t%d:%.9g" % (v, bl[v]))
            continue
        if st == 0:
            out.append("(")
            stack.append((v, 1)); stack.append((left[v], 0))
        elif st == 1:
            out.append(",")
            stack.append((v, 2)); stack.append((right[v], 0))
        else:
            out.append(")" if v == 0 else "):%.9g" % bl[v])
    out.append(";")
    return "".join(out)


def synth_problem(ntips, seed=1, numsites=10000, root_age=100.0, ncal=8):
    """Synthetic dating problem of SURVEY.md §8(d): Yule tree, fixed root age, `ncal` bounded
    interior calibrations (min < max around the true age, so barrier rows exist)."""
    nw = synth_yule_newick(ntips, seed=seed, root_age=root_age)
    parent, children, name, bl = parse_newick(nw)
    n = len(parent)
    # true ages are not stored in the Newick; recover relative ages from a clock-like proxy:
    # use the generator again for calibration placement instead -> choose clades by size
    rng = np.random.RandomState(seed + 7919)
    # descendant tip sample: first and last tip of the clade (their MRCA is the clade root)
    first_tip = [-1] * n; last_tip = [-1] * n; size = [0] * n; height = [0.0] * n
    for v in range(n - 1, -1, -1):
        if not children[v]:
            first_tip[v] = v; last_tip[v] = v; size[v] = 1
        else:
            first_tip[v] = first_tip[children[v][0]]; last_tip[v] = last_tip[children[v][-1]]
            size[v] = sum(size[c] for c in children[v])
    all_tips = [name[first_tip[0]], name[last_tip[0]]]
    cals = [(all_tips, root_age, root_age)]
    # time-depth proxy: clade size -> expected age under Yule ~ root_age * log(size)/log(ntips)
    cand = [v for v in range(1, n) if children[v] and size[v] >= max(4, ntips // 2000)]
    rng.shuffle(cand)
    used = 0
    for v in cand:
        if used >= ncal:
            break
        # keep calibrations nested-safe: loose windows well inside (0, root_age)
        est = root_age * math.log(size[v]) / math.log(ntips + 1)
        lo = max(0.5, 0.25 * est); hi = min(root_age * 0.98, 3.0 * est + 5.0)
        if lo < hi:
            cals.append(([name[first_tip[v]], name[last_tip[v]]], lo, hi))
            used += 1
    ft = flatten_newick(nw, numsites, cals, seed=seed)
    return ft, nw, cals
