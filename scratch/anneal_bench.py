"""GPU timing of the batched annealer (scratch; numbers copied into profiles/ by hand)."""
import json, sys, time
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from golden_util import Golden
from treepl_b200 import cv, engine as eng
out = {}
for name, B in (("M1155", 296), ("big_test", 1024)):
    flat = dict(Golden(name).flat)
    n = len(flat["parents"])
    plp = eng.PLCalcParallel.from_flat(flat, device=0)
    fp, P = eng.generate_param_order_vector(flat["free"], lf=False)
    plp.set_freeparams(P, False, fp)
    plp.batch_configure(B, lane_smoothing=np.full(B, 10.0))
    x0 = cv.start_vector(flat, P, n - 1)
    X0 = np.repeat(x0[:, None], B, axis=1)
    F0, _ = plp.calc_pl_grad_batch(X0[:, :2].copy() if False else X0, mode=eng.GRAD_ANALYTIC, want_grad=False)
    for rep in range(2):
        t = time.perf_counter()
        X, F, status, iters = plp.anneal_batch(X0, seed=3)
        wall = time.perf_counter() - t
    r = {"tree": name, "nodes": n, "chains": B, "iterations_per_chain": int(iters[0]), "wall_s": wall,
         "proposals_per_second": float(B * iters[0] / wall), "F0": float(F0[0]), "F_min": float(F.min()), "F_max": float(F.max())}
    out[name] = r
    print(json.dumps(r), flush=True)
    plp.close()
json.dump(out, open("gpurun_out/anneal_bench.json", "w"), indent=1)
