import sys, time, json
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from golden_util import Golden
from treepl_b200 import cv, treeflat as tf
gd = Golden("M1155")
cals = [(t, lo, hi) for t, lo, hi in gd.meta["calibrations"]]
ft = tf.flatten_newick(gd.meta["newick"], gd.meta["numsites"], cals, seed=1)
flat = dict(gd.flat)       # the reference's own start dates
tips = ft.tips_postorder
grid = [1000.0, 100.0, 10.0, 1.0, 0.1]
t = time.time()
res = cv.loocv(flat, grid, tips, verbose=False, max_iter=int(sys.argv[1]) if len(sys.argv) > 1 else 3000)
print("wall", time.time() - t)
print("chisq", res["chisq"])
print("ref  ", [2730.97, 2365.8, 2130.28, 2178.18, 2187.12])
print({k: v for k, v in res.items() if k not in ("terms",)})
