"""Result writers (treepl_b200/output.py) against the reference's own output of examples/test.cppr8s."""
import numpy as np

from golden_util import Golden
from treepl_b200 import output
from treepl_b200 import treeflat as tf

REF_DATED = ("(((a:3.997337,b:3.997337):4.002084,((c:3.207058,d:3.207058):1.600094,e:4.807152):3.192270):2.000579,"
             "(f:6.626589,g:6.626589):3.373411);")


def test_dated_tree_writer_reproduces_the_reference_file_format():
    gd = Golden("test")
    flat, names = gd.flat, gd.meta["names"]
    n = len(flat["parents"])
    # node dates of the reference's final tree, read off its own output: tips at 0, heights by summing branch lengths
    parent, children, name, bl = tf.parse_newick(REF_DATED)
    height = [0.0] * len(parent)
    for v in range(len(parent) - 1, -1, -1):
        if children[v]:
            height[v] = height[children[v][0]] + bl[children[v][0]]
    by_tips = {}
    def tipset(v):
        return frozenset([name[v]]) if not children[v] else frozenset().union(*[tipset(c) for c in children[v]])
    for v in range(len(parent)):
        by_tips[tipset(v)] = height[v]
    kids = output._children_lists(flat)
    def tips_of(i):
        return frozenset([names[i]]) if not kids[i] else frozenset().union(*[tips_of(c) for c in kids[i]])
    fp, P = tf.generate_param_order_vector(flat["free"], lf=False)
    x = np.zeros(P)
    for i in range(n):
        if fp[n + i] != -1:
            x[fp[n + i]] = by_tips[tips_of(i)]
    x[: n - 1] = 0.5
    got = output.dated_tree(flat, names, fp, x)
    # same topology, order, names and 6-decimal fixed format; lengths equal up to the rounding of re-summed heights
    strip = lambda s: "".join(ch for ch in s if not (ch.isdigit() or ch == "."))
    assert strip(got) == strip(REF_DATED)
    g = [float(t.split(")")[0].split(",")[0]) for t in got.split(":")[1:]]
    r = [float(t.split(")")[0].split(",")[0]) for t in REF_DATED.split(":")[1:]]
    assert np.allclose(g, r, atol=2.1e-6)
    rt = output.rate_tree(flat, names, fp, x, gd.meta["numsites"])
    assert rt.count(":%.6f" % (0.5 / gd.meta["numsites"])) == n - 1


def test_cv_out_lines_use_the_reference_stream_format():
    lines = output.cv_out_lines([1000.0, 100.0, 10.0, 1.0, 0.1], [2730.9716188, 2365.79757, 2130.2758, 2178.17067, 2187.12120])
    assert lines == ["chisq: (1000) 2730.97", "chisq: (100) 2365.8", "chisq: (10) 2130.28", "chisq: (1) 2178.17", "chisq: (0.1) 2187.12"]


def test_writer_is_iterative_on_a_deep_caterpillar():
    n_tips = 30000                     # depth 30,000: a recursive writer would overflow the stack
    nw = "(" * (n_tips - 1) + "t0:1" + "".join(",t%d:1):1" % k for k in range(1, n_tips - 1)) + ",t%d:1);" % (n_tips - 1)
    parent, children, name, bl = tf.parse_newick(nw)
    n = len(parent)
    flat = {"parents": np.array(parent), "child_counts": np.array([len(c) for c in children]),
            "children": np.array([c for ch in children for c in ch])}
    s = output.newick(flat, name, np.ones(n))
    assert s.count(",") == n_tips - 1 and s.endswith(";") and s.count("(") == n_tips - 1
