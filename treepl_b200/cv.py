"""Cross-validation over a smoothing grid (leave-one-out or random-subsample folds) as ONE batched device-resident
optimisation, sharded over the GPUs of a box (SURVEY.md §8 e, f1, f4; reference: the CV loop of src/main.cpp:588-822).

Reference flow: for each smoothing value (sequentially, warm-started from the previous one) fit the full data,
then for each tip prune it (CV mask), re-optimise from the full-data optimum, and score
chi^2 += (c_tip - r_par d_tip)^2 / (r_par d_tip)  (src/main.cpp:737-757), folds in an OpenMP loop.
Here every (smoothing value, fold) pair is one vector of the batch: per-vector smoothing + per-vector CV mask
(tpl_batch_configure), one launch of the PL kernels per trial point for all of them.
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import batchopt
from . import engine as eng


def start_vector(flat, P, n_rates):
    """A feasible start in the full PL layout: the reference's start dates (get_feasible_start_dates) and one
    clock rate sum(c) / sum(durations) on every branch (the LF phase's job in the reference)."""
    n = len(flat["parents"])
    par = flat["parents"]
    dates = flat["start_dates"]
    dur = dates[par[1:]] - dates[1:]
    rate = float(np.sum(flat["c"][1:]) / np.sum(dur))
    x = np.empty(P)
    x[:n_rates] = rate
    x[n_rates:] = dates[flat["free"] == 1]
    return x


def loocv_chisq(flat, X, lane_tip, freeparams):
    """chi-square term of each vector (src/main.cpp:737-757). X: [P,B] optimised vectors, lane_tip[b] = pruned tip."""
    n = len(flat["parents"])
    par = flat["parents"]
    fp = np.asarray(freeparams)
    out = np.zeros(X.shape[1])
    for b, tip in enumerate(lane_tip):
        p = par[tip]
        if p == 0:
            sib = [j for j in range(n) if par[j] == 0 and j != tip][-1]      # the loop keeps the LAST match
            r = X[fp[sib], b]
        else:
            r = X[fp[p], b]
        hp = X[fp[n + p], b] if fp[n + p] != -1 else flat["start_dates"][p]
        d = hp - flat["start_dates"][tip]
        if d == 0:
            d = 2.2250738585072014e-308
        expe = r * d
        c = flat["c"][tip]
        out[b] = 0.0 if (d < 1e-100 or expe < 1e-100) else (c - expe) ** 2 / expe
    return out


def _minimize(plp, X0, lo, hi, nr, dscale, max_iter, history, verbose, **opt):
    """-> (X, F, info) with info keys iterations (max accepted steps), nfev (batched evaluations), converged [B].
    method (option): 1 = batched truncated Newton (default; gradient-converged optima in far fewer evaluations),
    0 = batched L-BFGS (also serves the log-penalty objective)."""
    P, B = X0.shape
    opt.setdefault("method", 1)
    # chi-square is far more sensitive to the last digits of the optimum than the objective is (at small
    # smoothing the pruned tip's prediction hangs on weakly determined rates): converge tightly
    if opt["method"] == 0:
        opt.setdefault("ftol", 1e-15)
        opt.setdefault("patience", 40)
    X, F, info = plp.optimize_batch(X0, lo, hi, mode=eng.GRAD_AD, history=history, max_steps=20 * max_iter,
                                    date_scale=dscale, use_graph=1, **opt)
    bad = np.nonzero(info["status"] == 4)[0]
    if bad.size:
        raise ValueError("start point infeasible for vectors %s" % bad.tolist()[:8])
    return X, F, {"iterations": int(info["iterations"].max()), "nfev": info["steps"],
                  "converged": (info["status"] >= 1) & (info["status"] <= 3), "device_ms": info["device_ms"],
                  "launches": info["launches"], "cg_iterations": info["cg_iterations"]}


def _minimize_together(dist, dev, plp, *args, **opt):
    """_minimize on every rank; a failure on one rank (e.g. an infeasible start) is raised on all of them instead of
    leaving the others blocked in the next collective."""
    from . import cvshard

    err, res = None, None
    try:
        res = _minimize(plp, *args, **opt)
    except (ValueError, eng.EngineError) as ex:
        err = ex
    cvshard.raise_together(err, dist=dist, device=dev)
    return res


def sample_random_groups(flat, tips, n_groups, frac, seed=1):
    """Random-subsample folds (src/main.cpp:609-656): n_groups groups of round(frac * #tips) tips drawn without
    replacement among the tips that are not children of the root, no two of them sisters (the reference's
    check_possible_cv_nodes has an uninitialised index; the intent — never prune both children
    of a node — is what is implemented).  numpy RNG instead of libc rand(): groups are not the reference's."""
    par = flat["parents"]
    rng = np.random.RandomState(seed)
    size = int(len(tips) * frac + 0.5)
    poss_all = [int(t) for t in tips if par[t] != 0]
    groups = []
    for _ in range(n_groups):
        poss = list(poss_all)
        taken_parents = set()
        g = []
        fails = 0
        while len(g) < size and poss:
            k = int(rng.randint(len(poss)))
            t = poss[k]
            if par[t] in taken_parents:
                fails += 1
                if fails > 1000:
                    raise RuntimeError("could not sample a sister-free fold")
                continue
            g.append(t)
            taken_parents.add(par[t])
            poss.pop(k)
        groups.append(g)
    return groups


def group_chisq(flat, x, group, freeparams, randomcv):
    """chi-square contribution of one fold from its optimised vector x [P] (src/main.cpp:737-757 LOOCV,
    :769-793 random-subsample: mean over the group's tips).  Vectorised over the fold's tips (a random fold of
    the 100k-tip tree prunes 10,000 of them)."""
    n = len(flat["parents"])
    par = np.asarray(flat["parents"])
    fp = np.asarray(freeparams)
    tips = np.asarray(list(group), dtype=np.int64)
    p = par[tips]
    rslot = fp[np.maximum(p, 0)]
    r = np.where(rslot >= 0, x[np.maximum(rslot, 0)], 0.0)
    for k in np.flatnonzero(p == 0):                      # child of the root: the rate of its LAST sibling (main.cpp:741-746)
        sib = [j for j in np.flatnonzero(par == 0) if j != tips[k]][-1]
        r[k] = x[fp[sib]]
    dslot = fp[n + p]
    hp = np.where(dslot >= 0, x[np.maximum(dslot, 0)], flat["start_dates"][p])
    d = hp - flat["start_dates"][tips]
    d = np.where(d == 0, 2.2250738585072014e-308, d)
    expe = r * d
    c = flat["c"][tips]
    with np.errstate(divide="ignore", invalid="ignore"):
        terms = np.where((d < 1e-100) | (expe < 1e-100), 0.0, (c - expe) ** 2 / expe)
    total = 0.0
    for v in terms:                                        # tip order, fixed
        total += float(v)
    return total / (len(tips) if randomcv else 1.0)


def single_tip_chisq(flat, X, tips, freeparams):
    """group_chisq for one-tip folds, all lanes at once: X [P, B], tips[b] = the tip pruned in lane b
    (src/main.cpp:737-757; same arithmetic per lane as group_chisq)."""
    n = len(flat["parents"])
    par = np.asarray(flat["parents"])
    fp = np.asarray(freeparams)
    tips = np.asarray(tips, dtype=np.int64)
    lanes = np.arange(tips.shape[0])
    p = par[tips]
    rnode = p.copy()
    root_kids = np.flatnonzero(par == 0)
    for k in np.flatnonzero(p == 0):                      # child of the root: the rate of its LAST sibling (main.cpp:741-746)
        rnode[k] = [j for j in root_kids if j != tips[k]][-1]
    rslot = fp[rnode]
    r = np.where(rslot >= 0, X[np.maximum(rslot, 0), lanes], 0.0)
    dslot = fp[n + p]
    hp = np.where(dslot >= 0, X[np.maximum(dslot, 0), lanes], flat["start_dates"][p])
    d = hp - flat["start_dates"][tips]
    d = np.where(d == 0, 2.2250738585072014e-308, d)
    expe = r * d
    c = flat["c"][tips]
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where((d < 1e-100) | (expe < 1e-100), 0.0, (c - expe) ** 2 / expe)


def fold_probe_rows(flat, lane_groups, freeparams):
    """The X rows a fold's chi-square needs (src/main.cpp:737-793), per lane: for the j-th pruned tip of lane b,
    rows[2 j, b] = rate row of its parent (a child of the root: of its LAST sibling), rows[2 j + 1, b] = date row of its
    parent (-1: the parent's date is fixed).  -> (rows [2 L, B] int32, tips [L, B] int64, -1 = padding), L = longest group."""
    n = len(flat["parents"])
    par = np.asarray(flat["parents"])
    fp = np.asarray(freeparams)
    B = len(lane_groups)
    L = max((len(g) for g in lane_groups), default=0)
    tips = -np.ones((L, B), dtype=np.int64)
    for b, g in enumerate(lane_groups):
        tips[: len(g), b] = g
    valid = tips >= 0
    t = np.where(valid, tips, 0)
    p = par[t]
    rnode = p.copy()
    root_kids = np.flatnonzero(par == 0)
    for j, b in zip(*np.nonzero(valid & (p == 0))):       # child of the root: the rate of its LAST sibling (main.cpp:741-746)
        rnode[j, b] = [k for k in root_kids if k != tips[j, b]][-1]
    rows = -np.ones((2 * L, B), dtype=np.int32)
    rows[0::2] = np.where(valid, fp[rnode], -1)
    rows[1::2] = np.where(valid, fp[n + p], -1)
    return rows, tips


def chisq_from_probes(flat, probes, rows, tips, randomcv):
    """chi-square term of every lane from the probed rows (same arithmetic, tip by tip in group order, as group_chisq)."""
    par = np.asarray(flat["parents"])
    valid = tips >= 0
    t = np.where(valid, tips, 0)
    p = par[t]
    r = np.where(rows[0::2] >= 0, probes[0::2], 0.0)
    hp = np.where(rows[1::2] >= 0, probes[1::2], flat["start_dates"][p])
    d = hp - flat["start_dates"][t]
    d = np.where(d == 0, 2.2250738585072014e-308, d)
    expe = r * d
    c = flat["c"][t]
    with np.errstate(divide="ignore", invalid="ignore"):
        terms = np.where((d < 1e-100) | (expe < 1e-100), 0.0, (c - expe) ** 2 / expe)
    terms = np.where(valid, terms, 0.0)
    total = np.zeros(tips.shape[1])
    for j in range(tips.shape[0]):                         # tip order, fixed
        total = total + terms[j]
    cnt = valid.sum(axis=0)
    return total / np.maximum(cnt, 1) if randomcv else total


def cross_validate(flat, smoothing_grid, groups, randomcv=False, device=0, log_pen=False, max_iter=3000, verbose=False,
                   history=8, dist=None, **opt):
    """-> dict(chisq per smoothing value, best smoothing, timings).  groups: list of folds, each a list of tip node
    numbers (LOOCV: one tip per fold in externalNodes order, main.cpp:603-608).
    Every (smoothing value, fold) pair is one work item.  With `dist` (torch.distributed, one process per GPU) the
    items are dealt block-cyclically to the ranks (cvshard.shard_items), each rank optimises its share as one batch,
    and the per-item chi-square terms are all-gathered and summed in item order (deterministic for every world size).
    The optimiser is the library's own device-resident one (tpl_optimize_batch)."""
    from . import cvshard

    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    n = len(flat["parents"])
    plp = eng.PLCalcParallel.from_flat(flat, device=device)
    plp.set_log_pen(log_pen)
    fp, P = eng.generate_param_order_vector(flat["free"], lf=False)
    plp.set_freeparams(P, False, fp)
    nr = n - 1
    lo, hi = batchopt.reference_bounds(flat, fp, P)
    K, T = len(smoothing_grid), len(groups)
    t0 = time.perf_counter()
    # ---- full-data fits: one vector per smoothing value, dealt to the ranks; the optima are all-gathered ----
    dev = None
    if dist is not None and dist.get_backend() == "nccl":
        import torch

        dev = torch.device("cuda", device)
    mine_k = cvshard.shard_items(K, rank, world)
    B1 = max(2, len(mine_k) + (len(mine_k) % 2))
    lam1 = np.array([smoothing_grid[mine_k[min(b, len(mine_k) - 1)]] if mine_k else smoothing_grid[0] for b in range(B1)], dtype=np.float64)
    plp.batch_configure(B1, lane_smoothing=lam1)
    x0 = start_vector(flat, P, nr)
    par = flat["parents"]
    dscale = float(np.median(np.maximum(flat["start_dates"][par[1:]] - flat["start_dates"][1:], 1e-12)))
    Xl, Fl, info1 = _minimize_together(dist, dev, plp, np.repeat(x0[:, None], B1, axis=1), lo, hi, nr, dscale, max_iter, history, verbose, **opt)
    packed = {k: np.concatenate([Xl[:, b], [Fl[b], float(info1["converged"][b])]]) for b, k in enumerate(mine_k)}
    allfull = cvshard.gather_vectors(packed, K, P + 2, dist=dist, device=dev)
    Xfull, Ffull, full_conv = allfull[:P], allfull[P], allfull[P + 1]
    t1 = time.perf_counter()
    # ---- folds: one vector per (smoothing value, fold); this rank's share ----
    n_items = K * T
    mine = np.asarray(cvshard.shard_items(n_items, rank, world), dtype=np.int64)
    B2 = len(mine)
    B2p = max(2, B2 + (B2 % 2))
    lane_item = mine[np.minimum(np.arange(B2p), max(B2 - 1, 0))] if B2 else np.zeros(B2p, dtype=np.int64)   # padding repeats the last item
    lane_k, lane_g = lane_item // T, lane_item % T
    lam2 = np.asarray(smoothing_grid, dtype=np.float64)[lane_k]
    masks = [groups[g] for g in lane_g]
    plp.batch_configure(B2p, lane_smoothing=lam2, masks=masks)
    # every fold starts from the full-data optimum of its smoothing value (main.cpp:681,702-704): the K optima cross the bus
    # once and are expanded on the device; only the rows the chi-square needs come back (tpl_optimize_batch_cols)
    rows, ptips = fold_probe_rows(flat, masks, fp)
    if opt.get("method", 1) == 1:
        err, res2 = None, None
        try:
            # only the columns this rank's vectors start from (never more columns than vectors)
            used, col = np.unique(lane_k, return_inverse=True)
            Fcv, probes, _, oi = plp.optimize_batch_cols(np.ascontiguousarray(Xfull[:, used]), col.astype(np.int32), lo, hi, probe_rows=rows,
                                                         max_steps=20 * max_iter, date_scale=dscale, use_graph=1, **opt)
            bad = np.nonzero(oi["status"] == 4)[0]
            if bad.size:
                raise ValueError("start point infeasible for vectors %s" % bad.tolist()[:8])
            res2 = (Fcv, probes, oi)
        except (ValueError, eng.EngineError) as ex:
            err = ex
        cvshard.raise_together(err, dist=dist, device=dev)
        Fcv, probes, oi = res2
        info2 = {"iterations": int(oi["iterations"].max()), "nfev": oi["steps"], "converged": (oi["status"] >= 1) & (oi["status"] <= 3),
                 "device_ms": oi["device_ms"], "launches": oi["launches"], "cg_iterations": oi["cg_iterations"]}
    else:
        # method 0 (L-BFGS) has no start-column entry: the batch is expanded and read back on the host
        Xcv, Fcv, info2 = _minimize_together(dist, dev, plp, np.ascontiguousarray(Xfull[:, lane_k]), lo, hi, nr, dscale, max_iter, history,
                                             verbose, **opt)
        probes = np.where(rows >= 0, Xcv[np.maximum(rows, 0), np.arange(B2p)[None, :]], 0.0)
    t2 = time.perf_counter()
    local = chisq_from_probes(flat, probes, rows, ptips, randomcv)[:B2]
    full, _ = cvshard.gather_fold_scores(local, n_items, dist=dist, device=dev)
    terms = full.reshape(K, T)
    chisq = np.zeros(K)
    for k in range(K):
        for g in range(T):            # fold order, fixed
            chisq[k] += terms[k, g]
    t3 = time.perf_counter()
    plp.close()
    return {"smoothing": list(smoothing_grid), "chisq": chisq.tolist(), "best": float(smoothing_grid[int(np.argmin(chisq))]),
            "full_objective": Ffull[:K].tolist(), "full_converged": int(full_conv.sum()), "full_seconds": t1 - t0, "cv_seconds": t2 - t1, "gather_seconds": t3 - t2,
            "full_info": {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in info1.items()},
            "cv_iterations": info2["iterations"], "cv_nfev": info2["nfev"], "cv_converged": int(info2["converged"][:B2].sum()),
            "full_device_ms": info1.get("device_ms"), "cv_device_ms": info2.get("device_ms"),
            "launches": (info1.get("launches") or 0) + (info2.get("launches") or 0),
            "cg_iterations": (info1.get("cg_iterations") or 0) + (info2.get("cg_iterations") or 0),
            "n_vectors": n_items, "n_local_vectors": B2, "terms": terms}


def loocv(flat, smoothing_grid, tips, **kw):
    """Leave-one-out CV: one fold per tip (tips in externalNodes order, main.cpp:603-608)."""
    return cross_validate(flat, smoothing_grid, [[int(t)] for t in tips], randomcv=False, **kw)


def restart_starts(flat, P, ks, jitter=0.3, seed=1):
    """Start vectors [P, len(ks)] of the multi-start job: rates jittered log-normally around the clock-like start, feasible start
    dates.  Start k depends on (seed, k) only (one counter-based Philox stream per start, key = (seed, k)), so the set of starts
    is the same for every world size - and for the reference arm of the benchmark, which optimises the first few of them on the
    CPU.  Blocks of starts are drawn and exponentiated on a few threads (numpy releases the GIL in both): 1,024 starts of the
    9,313-node tree cost 0.75 s of the job's 1.4 s when drawn one Mersenne-Twister stream after the other."""
    from concurrent.futures import ThreadPoolExecutor

    nr = len(flat["parents"]) - 1
    x0 = start_vector(flat, P, nr)
    ks = list(ks)
    X0 = np.empty((P, len(ks)))
    X0[nr:] = x0[nr:, None]

    def block(lo):
        hi = min(len(ks), lo + 16)
        J = np.stack([np.random.Generator(np.random.Philox(key=[int(seed), int(k)])).standard_normal(nr) for k in ks[lo:hi]])
        X0[:nr, lo:hi] = x0[:nr, None] * np.exp(jitter * J).T

    starts = range(0, len(ks), 16)
    if len(ks) <= 32:
        for lo in starts:
            block(lo)
    else:
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
            list(ex.map(block, starts))
    return X0


def multistart(flat, smoothing, n_starts, jitter=0.3, seed=1, device=0, dist=None, log_pen=False, **opt):
    """n_starts independent optimisations from random starts (rates jittered log-normally around the clock-like
    start, feasible start dates), sharded over the ranks; the global best is exchanged with
    cvshard.best_of_restarts (all-gather of the best objectives, broadcast of the winner's vector).
    Start k depends on (seed, k) only, so the set of starts is the same for every world size.
    -> dict(best objective, best vector, per-start objectives of this rank, timings)."""
    from . import cvshard

    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    n = len(flat["parents"])
    plp = eng.PLCalcParallel.from_flat(flat, device=device)
    plp.set_log_pen(log_pen)
    fp, P = eng.generate_param_order_vector(flat["free"], lf=False)
    plp.set_freeparams(P, False, fp)
    plp.smoothing = float(smoothing)
    nr = n - 1
    lo, hi = batchopt.reference_bounds(flat, fp, P)
    mine = cvshard.shard_items(n_starts, rank, world)
    B = max(2, len(mine) + (len(mine) % 2))
    x0 = start_vector(flat, P, nr)
    ks = [mine[min(b, len(mine) - 1)] if mine else 0 for b in range(B)]
    X0 = restart_starts(flat, P, ks, jitter, seed)
    plp.batch_configure(B)
    par = flat["parents"]
    dscale = float(np.median(np.maximum(flat["start_dates"][par[1:]] - flat["start_dates"][1:], 1e-12)))
    t0 = time.perf_counter()
    dev = None
    if dist is not None and dist.get_backend() == "nccl":
        import torch

        dev = torch.device("cuda", device)
    X, F, info = _minimize_together(dist, dev, plp, X0, lo, hi, nr, dscale, 3000, 8, False, **opt)
    t1 = time.perf_counter()
    Fm = F[: len(mine)]
    ok = info["converged"][: len(mine)]
    have = bool(len(mine)) and bool(ok.any())         # a rank whose starts all failed submits +inf, not an unconverged value
    kbest = int(np.argmin(np.where(ok, Fm, np.inf))) if have else 0
    fbest, xbest, winner = cvshard.best_of_restarts(Fm[kbest] if have else np.inf, X[:, kbest], dist=dist, device=dev)
    plp.close()
    return {"best": fbest, "x": xbest, "winner_rank": winner, "local_objectives": Fm, "local_converged": int(ok.sum()),
            "n_local": len(mine), "optimize_seconds": t1 - t0, "nfev": info["nfev"], "cg_iterations": info.get("cg_iterations"),
            "launches": info.get("launches"), "device_ms": info.get("device_ms")}
