"""GPU measurements of the native optimiser (scratch; results are copied into profiles/ by hand)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from golden_util import Golden
from treepl_b200 import batchopt, cv, engine as eng, treeflat as tf

out = {}
what = sys.argv[1:] or ["pcie", "m1155", "big"]

if "pcie" in what:
    import torch
    n = 614 * 1000 * 1000 // 8
    h = torch.empty(n, dtype=torch.float64).pin_memory()
    h2 = torch.empty(n, dtype=torch.float64).pin_memory()
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    d2 = torch.empty(n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def timed(fn, reps=5):
        fn(); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t) / reps
    t_h2d = timed(lambda: d.copy_(h, non_blocking=True))
    t_d2h = timed(lambda: h2.copy_(d2, non_blocking=True))
    def both():
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    t_both = timed(both)
    out["pcie"] = {"bytes": n * 8, "h2d_GBs": n * 8 / t_h2d / 1e9, "d2h_GBs": n * 8 / t_d2h / 1e9,
                   "duplex_each_GBs": n * 8 / t_both / 1e9}
    print(out["pcie"], flush=True)
    del h, h2, d, d2

if "m1155" in what:
    gd = Golden("M1155")
    cals = [(t, lo, hi) for t, lo, hi in gd.meta["calibrations"]]
    ft = tf.flatten_newick(gd.meta["newick"], gd.meta["numsites"], cals, seed=1)
    grid = [1000.0, 100.0, 10.0, 1.0, 0.1]
    for hist in (8, 16):
        for rep in range(2):
            t = time.perf_counter()
            res = cv.loocv(dict(gd.flat), grid, ft.tips_postorder, max_iter=6000, native=True, history=hist)
            wall = time.perf_counter() - t
        r = {k: v for k, v in res.items() if k not in ("terms", "full_info")}
        r["wall_s"] = wall
        out["m1155_h%d" % hist] = r
        print(hist, json.dumps(r), flush=True)

if "tol" in what:
    gd = Golden("M1155")
    cals = [(t, lo, hi) for t, lo, hi in gd.meta["calibrations"]]
    ft = tf.flatten_newick(gd.meta["newick"], gd.meta["numsites"], cals, seed=1)
    grid = [1000.0, 100.0, 10.0, 1.0, 0.1]
    for hist in (8, 16):
        for ftol, pat in ((1e-12, 5), (1e-13, 10), (1e-14, 20), (1e-15, 40)):
            t = time.perf_counter()
            res = cv.loocv(dict(gd.flat), grid, ft.tips_postorder, max_iter=20000, native=True, history=hist, ftol=ftol, patience=pat)
            wall = time.perf_counter() - t
            print(hist, ftol, pat, "wall %.3f" % wall, "nfev", res["cv_nfev"], "conv", res["cv_converged"], res["chisq"], res["full_objective"], flush=True)

if "big" in what:
    import bench
    flat = bench.build_problem(100000)
    fp, P = tf.generate_param_order_vector(flat["free"], lf=False)
    n = flat["parents"].shape[0]
    lo, hi = batchopt.reference_bounds(flat, fp, P)
    par = flat["parents"]
    ds = float(np.median(np.maximum(flat["start_dates"][par[1:]] - flat["start_dates"][1:], 1e-12)))
    for B, hist in ((64, 8), (160, 8), (160, 4)):
        plp = eng.PLCalcParallel.from_flat(flat, device=0)
        plp.set_freeparams(P, False, fp)
        lam = np.repeat(10.0 ** np.linspace(3, -1, 16), (B + 15) // 16)[:B]
        plp.batch_configure(B, lane_smoothing=lam)
        X0 = bench.make_vectors(flat, P, B, 7)
        for steps in (64, 256):
            t = time.perf_counter()
            X, F, info = plp.optimize_batch(X0, lo, hi, date_scale=ds, history=hist, max_steps=steps, check_every=32, use_graph=1)
            wall = time.perf_counter() - t
            F0, _ = plp.calc_pl_grad_batch(X0[:, :2].copy(), mode=eng.GRAD_AD) if False else (None, None)
            r = {"B": B, "history": hist, "steps": info["steps"], "device_ms": info["device_ms"],
                 "ms_per_step": info["device_ms"] / info["steps"], "wall_s": wall,
                 "iters_min_max": [int(info["iterations"].min()), int(info["iterations"].max())],
                 "F_first": float(F[0]), "F_last": float(F[-1]), "status": np.bincount(info["status"], minlength=6).tolist()}
            out["big_B%d_h%d_s%d" % (B, hist, steps)] = r
            print(json.dumps(r), flush=True)
        plp.close()

if "conv" in what:
    import bench
    flat = bench.build_problem(100000)
    fp, P = tf.generate_param_order_vector(flat["free"], lf=False)
    n = flat["parents"].shape[0]
    lo, hi = batchopt.reference_bounds(flat, fp, P)
    par = flat["parents"]
    ds = float(np.median(np.maximum(flat["start_dates"][par[1:]] - flat["start_dates"][1:], 1e-12)))
    B = 16
    lam = 10.0 ** np.linspace(3, -1, 16)
    x0 = cv.start_vector(flat, P, n - 1)
    X0 = np.repeat(x0[:, None], B, axis=1)
    for hist in (8,):
        plp = eng.PLCalcParallel.from_flat(flat, device=0)
        plp.set_freeparams(P, False, fp)
        plp.batch_configure(B, lane_smoothing=lam)
        Xc = X0
        tot = 0
        for rnd in range(2):
            t = time.perf_counter()
            X, F, info = plp.optimize_batch(Xc, lo, hi, date_scale=ds, history=hist, max_steps=5000, check_every=50, use_graph=1)
            wall = time.perf_counter() - t
            tot += info["steps"]
            print(json.dumps({"history": hist, "round": rnd, "steps": info["steps"], "total_steps": tot, "wall_s": wall, "ms_per_step": info["device_ms"] / info["steps"],
                              "status": np.bincount(info["status"], minlength=6).tolist(), "iters": [int(info["iterations"].min()), int(info["iterations"].max())],
                              "mean_log10_step": info["mean_log10_step"], "F": [float(F[0]), float(F[5]), float(F[10]), float(F[15])]}), flush=True)
            Xc = X
            if (info["status"] > 0).all():
                break
        plp.close()

if "tn" in what:
    gd = Golden("M1155")
    cals = [(t, lo, hi) for t, lo, hi in gd.meta["calibrations"]]
    ft = tf.flatten_newick(gd.meta["newick"], gd.meta["numsites"], cals, seed=1)
    grid = [1000.0, 100.0, 10.0, 1.0, 0.1]
    for method in (1, 0):
        for rep in range(2):
            t = time.perf_counter()
            res = cv.loocv(dict(gd.flat), grid, ft.tips_postorder, max_iter=6000, method=method)
            wall = time.perf_counter() - t
        r = {k: v for k, v in res.items() if k not in ("terms", "full_info")}
        r["wall_s"] = wall
        r["full_nfev"] = res["full_info"]["nfev"]
        out["m1155_method%d" % method] = r
        print(method, json.dumps(r), flush=True)
    import bench
    flat = bench.build_problem(100000)
    fp, P = tf.generate_param_order_vector(flat["free"], lf=False)
    n = flat["parents"].shape[0]
    lo, hi = batchopt.reference_bounds(flat, fp, P)
    par = flat["parents"]
    ds = float(np.median(np.maximum(flat["start_dates"][par[1:]] - flat["start_dates"][1:], 1e-12)))
    B = 16
    lam = 10.0 ** np.linspace(3, -1, 16)
    x0 = cv.start_vector(flat, P, n - 1)
    X0 = np.repeat(x0[:, None], B, axis=1)
    plp = eng.PLCalcParallel.from_flat(flat, device=0)
    plp.set_freeparams(P, False, fp)
    plp.batch_configure(B, lane_smoothing=lam)
    Xc = X0
    tot = 0
    for rnd in range(6):
        t = time.perf_counter()
        X, F, info = plp.optimize_batch(Xc, lo, hi, date_scale=ds, method=1, max_cg=int(sys.argv[2]) if len(sys.argv) > 2 else 200,
                                        max_steps=4000, check_every=50, use_graph=1)
        wall = time.perf_counter() - t
        tot += info["steps"]
        r = {"round": rnd, "steps": info["steps"], "total_steps": tot, "wall_s": wall, "ms_per_step": info["device_ms"] / info["steps"],
             "status": np.bincount(info["status"], minlength=6).tolist(), "newton_iters": [int(info["iterations"].min()), int(info["iterations"].max())],
             "cg": info["cg_iterations"], "mean_log10_step": info["mean_log10_step"], "F": [float(F[0]), float(F[5]), float(F[10]), float(F[15])]}
        out["tn100k_round%d" % rnd] = r
        print(json.dumps(r), flush=True)
        Xc = X
        if (info["status"] > 0).all():
            break
    plp.close()

if "restarts" in what:
    flat = dict(Golden("big_test").flat)
    fp, P = tf.generate_param_order_vector(flat["free"], lf=False)
    n = len(flat["parents"])
    lo, hi = batchopt.reference_bounds(flat, fp, P)
    par = flat["parents"]
    ds = float(np.median(np.maximum(flat["start_dates"][par[1:]] - flat["start_dates"][1:], 1e-12)))
    x0 = cv.start_vector(flat, P, n - 1)
    for B in (64, 1024):
        rng = np.random.RandomState(1)
        X0 = np.repeat(x0[:, None], B, axis=1)
        X0[: n - 1] *= np.exp(0.3 * rng.randn(n - 1, B))
        plp = eng.PLCalcParallel.from_flat(flat, device=0)
        plp.set_freeparams(P, False, fp)
        plp.smoothing = 1.0
        plp.batch_configure(B)
        for method in (1,):
            t = time.perf_counter()
            X, F, info = plp.optimize_batch(X0, lo, hi, date_scale=ds, method=method, max_steps=20000, check_every=50, use_graph=1)
            wall = time.perf_counter() - t
            r = {"B": B, "method": method, "steps": info["steps"], "wall_s": wall, "device_ms": info["device_ms"], "status": np.bincount(info["status"], minlength=6).tolist(),
                 "iters": [int(info["iterations"].min()), int(info["iterations"].max())], "cg": info["cg_iterations"], "Fmin": float(F.min()), "Fmax": float(F.max()),
                 "Fspread_rel": float((F.max() - F.min()) / abs(F.min()))}
            out["restarts_B%d_m%d" % (B, method)] = r
            print(json.dumps(r), flush=True)
        plp.close()

if "cv100k" in what:
    import bench
    flat = bench.build_problem(100000)
    n = flat["parents"].shape[0]
    tips = np.flatnonzero(flat["child_counts"] == 0)
    grid = list(10.0 ** np.linspace(3, -1, 16))
    groups = cv.sample_random_groups(flat, tips, n_groups=10, frac=0.1, seed=1)
    print("folds", len(groups), "tips per fold", len(groups[0]), flush=True)
    t = time.perf_counter()
    res = cv.cross_validate(flat, grid, groups, randomcv=True, max_iter=int(sys.argv[2]) if len(sys.argv) > 2 else 400, check_every=25)
    wall = time.perf_counter() - t
    r = {k: v for k, v in res.items() if k not in ("terms", "full_info")}
    r["wall_s"] = wall
    r["full_converged"] = int(np.sum(res["full_info"]["converged"]))
    r["full_nfev"] = res["full_info"]["nfev"]
    out["cv100k"] = r
    print(json.dumps(r), flush=True)

if "ncubig" in what or "ncusmall" in what:
    import bench
    if "ncubig" in what:
        flat = bench.build_problem(100000); B = 64
    else:
        flat = dict(Golden("M1155").flat); B = 296
    fp, P = tf.generate_param_order_vector(flat["free"], lf=False)
    lo, hi = batchopt.reference_bounds(flat, fp, P)
    par = flat["parents"]
    n = len(par)
    ds = float(np.median(np.maximum(flat["start_dates"][par[1:]] - flat["start_dates"][1:], 1e-12)))
    plp = eng.PLCalcParallel.from_flat(flat, device=0)
    plp.set_freeparams(P, False, fp)
    plp.batch_configure(B, lane_smoothing=np.full(B, 10.0))
    X0 = bench.make_vectors(flat, P, B, 7) if "ncubig" in what else np.repeat(cv.start_vector(flat, P, n - 1)[:, None], B, axis=1)
    X, F, info = plp.optimize_batch(X0, lo, hi, date_scale=ds, history=8, max_steps=64 if "tnm" not in what else 200, check_every=32, use_graph=0,
                                    method=1 if "tnm" in what else 0)
    print(info["steps"], info["device_ms"], F[:2])

json.dump(out, open("gpurun_out/opt_bench_%s.json" % "_".join(what), "w"), indent=1, default=str)
