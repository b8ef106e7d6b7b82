"""User-visible result files of a dating run, in the reference's formats:
  * the dated tree  `<outfile>`      — Newick whose branch lengths are the optimised durations   (src/main.cpp:866-873)
  * the rate tree   `<outfile>.r8s`  — branch lengths = rate / numsites                          (src/main.cpp:875-879)
  * the CV table    `cv.out`         — one line `chisq: (<smoothing>) <chi-square>` per value    (src/main.cpp:803-804)
The writer follows Node::getNewick(true) (src/node.cpp:154-175): children in Newick order, every non-root branch length in
std::fixed with the stream's default precision (6 decimals), names after the closing parenthesis.  Iterative, so a
100,000-tip caterpillar does not overflow the Python stack (the reference recurses)."""
from __future__ import annotations

import numpy as np


def _children_lists(flat):
    n = len(flat["parents"])
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum(flat["child_counts"])
    ch = np.asarray(flat["children"])
    return [ch[off[i]:off[i + 1]].tolist() for i in range(n)]


def newick(flat, names, lengths):
    """flat: the reference's flat vectors (children_vec in Newick order); names[i]: label of node i ('' = none);
    lengths[i]: branch length above node i.  -> Newick string ending in ';'."""
    kids = _children_lists(flat)
    out = []
    # explicit stack of (node, next child position)
    stack = [(0, 0)]
    while stack:
        node, k = stack.pop()
        ch = kids[node]
        if k == 0 and ch:
            out.append("(")
        if k < len(ch):
            if k > 0:
                out.append(",")
            stack.append((node, k + 1))
            stack.append((ch[k], 0))
            continue
        if ch:
            out.append(")")
        if names[node]:
            out.append(names[node])
        if node != 0:
            out.append(":%.6f" % lengths[node])
    return "".join(out) + ";"


def node_state(flat, freeparams, x):
    """-> (rates[N], dates[N], durations[N]) of one optimised parameter vector (what main() reads back from the engine,
    src/main.cpp:865-878): fixed dates keep their start values, durations[0] stays as given."""
    n = len(flat["parents"])
    fp = np.asarray(freeparams)
    par = np.asarray(flat["parents"])
    rates = np.asarray(flat["start_rates"], dtype=np.float64).copy()
    dates = np.asarray(flat["start_dates"], dtype=np.float64).copy()
    rs, ds = fp[:n], fp[n:]
    rates[rs >= 0] = x[rs[rs >= 0]]
    dates[ds >= 0] = x[ds[ds >= 0]]
    dur = np.asarray(flat["start_durations"], dtype=np.float64).copy()
    dur[1:] = dates[par[1:]] - dates[1:]
    return rates, dates, dur


def dated_tree(flat, names, freeparams, x):
    rates, dates, dur = node_state(flat, freeparams, x)
    return newick(flat, names, dur)


def rate_tree(flat, names, freeparams, x, numsites):
    rates, dates, dur = node_state(flat, freeparams, x)
    return newick(flat, names, rates / float(numsites))


def cv_out_lines(smoothing_grid, chisq):
    """the lines the reference appends to cv.out, default ostream formatting (6 significant digits)"""
    return ["chisq: (%g) %g" % (lam, c) for lam, c in zip(smoothing_grid, chisq)]


def write_results(outfile, flat, names, freeparams, x, numsites, cv_table=None, cvoutfile="cv.out"):
    with open(outfile, "w") as fh:
        fh.write(dated_tree(flat, names, freeparams, x) + "\n")
    with open(outfile + ".r8s", "w") as fh:
        fh.write(rate_tree(flat, names, freeparams, x, numsites) + "\n")
    if cv_table is not None:
        with open(cvoutfile, "w") as fh:
            fh.write("\n".join(cv_out_lines(*cv_table)) + "\n")
