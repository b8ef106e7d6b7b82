#!/usr/bin/env python
"""bench.py — PL objective+gradient evaluations per second (fp64) on a synthetic 100k-tip tree.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--tips T] [--impl reference]

One "step" = one pass of the hot path over one batch of B parameter vectors (one launch of the
batched objective+gradient kernel pair) on the workload BASELINE.json's metric is quoted on:
a 100,000-tip birth-death tree (N = 199,999 nodes, P = 299,996 parameters).

  value      evaluations/s with X, F, G resident in HBM (CUDA events on the launching stream)
  e2e        the same through the C-ABI batched entry with pinned HOST buffers: H2D of X and
             D2H of F and G inside the timed region
  roofline   main kernel: algorithmic bytes (16 P + 25 N / B per evaluation)
             / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own calc_pl + calc_gradient (oracle/_ref, built from the
             reference sources) on the host cores, bounded sample

N > 1 (torchrun): one process per GPU, every rank evaluates its own B vectors of the same tree
(CV folds / restarts are independent: weak scaling, no data-path collective); barrier + device
timing, max over ranks.  `--impl reference` times the reference CPU path alone (rank 0 only).

The default line also carries, at EVERY N, `sharded`: the optimiser-level jobs north_star shards over the GPUs (M1155 LOOCV,
1,024 restarts of big_test, LOOCV of the big_test tree = 23,285 optimisations, random-subsample CV of the 100k-tip tree), each
as wall seconds of the same job on N GPUs (strong scaling) with converged counts and chi-square tables; `--impl reference`
carries the reference's own optimiser stack on the host cores for them (whole job for M1155, bounded + extrapolated samples
for the big_test jobs).

Each of these jobs can also be the measured workload (BASELINE.json: "LOO-CV wall time at 1/8 B200"):
    python bench.py --workload loocv [--gpus N] [--steps K] [--impl reference]
one step = the complete leave-one-out cross-validation of examples/M1155.cppr8s (59 folds x 5 smoothing values,
tests/golden/M1155.npz) with the library's device-resident batched truncated Newton; the (smoothing, fold) vectors are
dealt to the ranks and the chi-square terms all-gathered over NCCL (strong scaling).  The reference arm replays
the reference's own CV loop around its optimize_full (annealer + TNC / NLopt) on the host cores (oracle/_ref).
    python bench.py --workload loocv_big [--gpus N]
the same for the 4,657-tip examples/big_test.cppr8s tree (23,285 optimisations of 13,967 parameters).
    python bench.py --workload restarts | restarts_cfg4 [--gpus N]
one step = 1,024 random-restart optimisations of the examples/big_test.cppr8s tree (BASELINE.json config 4), starts
dealt to the ranks, best-of-restarts exchange over NCCL (restarts_cfg4: the control file's smooth = 0.0001 + log_pen).
    python bench.py --workload cv100k [--gpus N] [--cv-steps S]
random-subsample CV over 16 smoothing values on the synthetic 100k-tip tree (BASELINE.json config 5).
    python bench.py --workload sweep
B sweep of the evaluation path and the B = 1 callback latency next to the reference's per-evaluation time.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL_DEBUG=VERSION makes NCCL print its banner on stdout, in front of the one JSON line this script owes the driver
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

METRIC = "pl_obj_grad_evals_per_sec"
UNIT = "evals/s"


def build_problem(tips, seed=1):
    from treepl_b200 import treeflat as tf
    cache = os.path.join(tempfile.gettempdir(), "treepl_b200_synth_%d_%d.npz" % (tips, seed))
    if os.path.exists(cache):
        z = np.load(cache)
        return {k: z[k] for k in z.files}
    ft, _, _ = tf.synth_problem(tips, seed=seed, ncal=16)
    d = ft.as_dict()
    try:        # atomic: the ranks of a multi-GPU run build the same tree side by side and must never read a half-written cache
        tmp = "%s.%d.tmp.npz" % (cache, os.getpid())
        np.savez(tmp, **d)
        os.replace(tmp, cache)
    except OSError:
        pass
    return d


def make_vectors(flat, P, B, seed, numsites=10000):
    """B feasible parameter vectors, [P, B] batch innermost: jittered rates, feasible start dates."""
    n = flat["parents"].shape[0]
    rng = np.random.RandomState(seed)
    free = flat["free"] == 1
    x_dates = flat["start_dates"][free]
    base = flat["start_rates"][1] * numsites
    X = np.empty((P, B), dtype=np.float64)
    X[: n - 1, :] = base * (1.0 + 0.1 * np.sin(np.arange(1, n)[:, None] + np.arange(B)[None, :]))
    X[n - 1:, :] = x_dates[:, None] * (1.0 - 1e-6 * rng.rand(1, B))
    return X


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [t.strip() for t in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(tips, batch, mode):
    """DRAM bytes per launch of the main kernel from the committed ncu capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        d = json.load(open(p))
        w = d["workload"]
        if w["tips"] == tips and w["batch"] == batch and w["mode"] == mode:
            return d["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def oracle_parity(flat, Xh, Fh, G, ad, lanes):
    """CHECKER, never timed: vectors `lanes` of the bench batch against oracle/pl_oracle.c (the plain-C restatement of
    pl_calc_parallel.cpp); -> max relative objective error and max gradient error (norm-wise floor of SURVEY 7)."""
    from oracle import plo
    o = plo.OracleProblem(flat)
    o.set_smoothing(1.0)
    o.set_freeparams(False)
    f_rel = g_rel = 0.0
    for b in lanes:
        xb = np.ascontiguousarray(Xh[:, b])
        f, g = o.calc_function_gradient(xb) if ad else o.calc_pl_and_gradient(xb)
        gb = G[:, b].cpu().numpy()
        f_rel = max(f_rel, abs(float(Fh[b]) - f) / max(abs(f), 1.0))
        scale = np.maximum(np.abs(g), 1e-8 * np.max(np.abs(g)))
        g_rel = max(g_rel, float(np.max(np.abs(gb - g) / scale)))
    return {"f_rel": f_rel, "g_rel": g_rel, "vectors": list(lanes), "checker": "oracle/pl_oracle.c",
            "gates": {"f_rel": 1e-10, "g_rel": 1e-8}}


def cpu_reference_rate(flat, x, seconds=12.0, nthreads=None):
    """Reference calc_pl + calc_gradient on the host cores -> (evals/s, threads, sample description)."""
    from oracle import refbind as rb
    if not rb.available():
        return None
    nthreads = nthreads or max(1, (os.cpu_count() or 1))
    rp = rb.RefProblem.from_flat(flat)
    rp.set_smoothing(1.0)
    rp.set_freeparams(False)
    t1 = rp.time_evals(x, 1, nthreads, True)
    reps = int(max(2, min(2000, seconds / max(t1, 1e-4))))
    t = rp.time_evals(x, reps, nthreads, True)
    rate = reps * nthreads / t
    rp.close()
    return rate, nthreads, "%d calc_pl+calc_gradient calls on each of %d threads (%.1f s)" % (reps, nthreads, t)


def run_reference_arm(args, flat, P):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    X = make_vectors(flat, P, 1, 1234)
    x = np.ascontiguousarray(X[:, 0])
    n = flat["parents"].shape[0]
    steps_total = args.steps + args.warmup
    budget = min(20.0, 150.0 / max(1, steps_total))
    res = None
    times = []
    from oracle import refbind as rb
    if not rb.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libtreepl_ref.so is not built"}))
        return
    nthreads = max(1, os.cpu_count() or 1)
    rp = rb.RefProblem.from_flat(flat)
    rp.set_smoothing(1.0)
    rp.set_freeparams(False)
    t1 = rp.time_evals(x, 1, nthreads, True)
    reps = int(max(1, min(500, budget / max(t1, 1e-4))))
    for s in range(steps_total):
        t = rp.time_evals(x, reps, nthreads, True)
        if s >= args.warmup:
            times.append(t)
    rp.close()
    tsum = sum(times)
    value = reps * nthreads * len(times) / tsum
    sample = "%d calc_pl+calc_gradient calls per thread per step, %d threads" % (reps, nthreads)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tsum / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic %d-tip Yule tree, N=%d nodes, P=%d params, PL (lf=false), smoothing=1"
                                   % (args.tips, n, P)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not args.no_sharded:
        # the sharded optimiser-level jobs of the b200 arm's `sharded` key: the reference's own optimiser stack on all host
        # threads - the whole M1155 LOOCV, and BOUNDED SAMPLES (extrapolated, labelled) of the two big_test jobs
        sh = {}
        for name, fn in (("loocv_M1155", ref_loocv_m1155), ("restarts", ref_sample_restarts), ("loocv_big_test", ref_sample_loocv_big)):
            try:
                sh[name] = fn(nthreads)
            except Exception as ex:
                sh[name] = {"error": repr(ex)}
        line["sharded"] = sh
    print(json.dumps(line))


M1155_GRID = [1000.0, 100.0, 10.0, 1.0, 0.1]
M1155_REF_CHISQ = [2730.97, 2365.8, 2130.28, 2178.18, 2187.12]      # reference run, seed 1, nthreads 1


def load_m1155():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import Golden
    from treepl_b200 import treeflat as tf
    gd = Golden("M1155")
    cals = [(t, lo, hi) for t, lo, hi in gd.meta["calibrations"]]
    ft = tf.flatten_newick(gd.meta["newick"], gd.meta["numsites"], cals, seed=1)
    return gd, cals, [int(t) for t in ft.tips_postorder]


BIG_GRID = [1000.0, 100.0, 10.0, 1.0, 0.1]          # the reference's default CV grid (cvstart 1000, cvstop 0.1, cvmultstep 0.1)
N_STARTS = 1024


def load_big_test():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import Golden
    return dict(Golden("big_test").flat)


def restart_vectors(flat, P, ks, jitter=0.3, seed=1):
    """the start vectors of cv.multistart, [len(ks), P] (start k depends on (seed, k) only)"""
    from treepl_b200 import cv
    return np.ascontiguousarray(cv.restart_starts(flat, P, ks, jitter, seed).T)


class Ranks:
    """torch.distributed plumbing shared by the sharded jobs: barrier + device sync around a wall-clock region,
    max over ranks."""

    def __init__(self, dist, local):
        import torch
        self.torch, self.dist, self.local = torch, dist, local
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.dev = torch.device("cuda", local)

    def fence(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn):
        self.fence()
        t = time.perf_counter()
        res = fn()
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        dt = time.perf_counter() - t
        return res, self.vmax(dt)

    def vmax(self, v):
        if self.dist is None:
            return float(v)
        tt = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX)
        return float(tt.item())

    def vsum(self, v):
        if self.dist is None:
            return v
        tt = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(tt)
        return float(tt.item())


def job_loocv(rk, which, runs=1, warm=0):
    """Complete leave-one-out cross-validation (main.cpp:588-822) of `which` = "M1155" (examples/M1155.cppr8s: 59 folds x
    5 smoothing values, P = 170) or "big_test" (the examples/big_test.cppr8s tree: 4,657 folds x 5 smoothing values =
    23,285 optimisations of P = 13,967); (smoothing, fold) vectors dealt to the ranks, chi-square terms all-gathered."""
    from treepl_b200 import cv
    if which == "M1155":
        gd, cals, tips = load_m1155()
        flat, grid = dict(gd.flat), M1155_GRID
    else:
        flat, grid = load_big_test(), BIG_GRID
        tips = [int(t) for t in np.flatnonzero(flat["child_counts"] == 0)]
    P = int((flat["free"] == 1).sum()) + len(flat["parents"]) - 1
    times, res = [], None
    for s in range(warm + runs):
        res, dt = rk.timed(lambda: cv.loocv(dict(flat), grid, tips, device=rk.local, max_iter=6000, dist=rk.dist))
        if s >= warm:
            times.append(dt)
    val = sum(times) / len(times)
    out = {"workload": "LOOCV of the %s tree: %d folds x %d smoothing values = %d optimisations of P=%d parameters" % (
               "examples/M1155.cppr8s" if which == "M1155" else "examples/big_test.cppr8s (comm_mol_rr.tre, N=9313)",
               len(tips), len(grid), len(tips) * len(grid), P),
           "wall_s": val, "runs": len(times), "n_gpus": rk.world, "chisq": res["chisq"], "best_smoothing": res["best"],
           "full_objective": res["full_objective"], "full_fits_converged": res["full_converged"],
           "fold_vectors_converged": int(rk.vsum(res["cv_converged"])), "fold_vectors": len(tips) * len(grid),
           "seconds": {"full_fits": res["full_seconds"], "folds": res["cv_seconds"], "scoring_and_gather": res["gather_seconds"]},
           "device_ms": {"full_fits": res["full_device_ms"], "folds": res["cv_device_ms"]},
           "batched_evaluations": {"full_fits": res["full_info"]["nfev"], "folds": res["cv_nfev"]},
           "hessian_vector_products_rank0": res["cg_iterations"], "gpu_launches": int(res["launches"]) * rk.world,
           "n_local_vectors": res["n_local_vectors"], "P": P}
    if which == "M1155":
        out["chisq_reference_run"] = M1155_REF_CHISQ
        out["chisq_max_rel_diff"] = max(abs(a - b) / b for a, b in zip(res["chisq"], M1155_REF_CHISQ))
    return out


def job_restarts(rk, runs=1, warm=0, config4=False):
    """BASELINE.json config 4: 1,024 batched random-restart optimisations of the examples/big_test.cppr8s tree
    (N = 9,313, P = 13,967), starts dealt to the ranks, best-of-restarts exchange.  config4: the control file's own
    objective (smooth = 0.0001, log_pen; big_test.cppr8s:6,14) instead of smoothing 1."""
    from treepl_b200 import cv
    flat = load_big_test()
    lam, lp = (0.0001, True) if config4 else (1.0, False)
    times, res = [], None
    for s in range(warm + runs):
        res, dt = rk.timed(lambda: cv.multistart(flat, lam, N_STARTS, device=rk.local, dist=rk.dist, check_every=10, log_pen=lp))
        if s >= warm:
            times.append(dt)
    val = sum(times) / len(times)
    return {"workload": "examples/big_test.cppr8s tree (N=9313, P=13967), smoothing %g%s: %d random-restart optimisations to |g| <= 1e-8 |f|"
                        % (lam, ", log_pen" if lp else "", N_STARTS),
            "wall_s": val, "runs": len(times), "n_gpus": rk.world, "optimisations_per_second": N_STARTS / val,
            "converged": int(rk.vsum(res["local_converged"])), "best_objective": res["best"], "winner_rank": res["winner_rank"],
            "objective_spread_rank0": [float(np.min(res["local_objectives"])), float(np.max(res["local_objectives"]))],
            "gpu_launches": int(res["launches"]) * rk.world, "batched_evaluations": res["nfev"],
            "hessian_vector_products_rank0": res["cg_iterations"], "n_local": res["n_local"],
            "seconds": {"optimize_rank0": res["optimize_seconds"], "device_rank0": 1e-3 * (res["device_ms"] or 0.0)}}


def job_cv100k(rk, tips_n, cv_steps, warm=0):
    """BASELINE.json config 5: random-subsample cross-validation (10 folds of 10 % of the tips, main.cpp:609-656) over 16
    smoothing values 10^3 ... 10^-1 on the synthetic 100,000-tip tree, (smoothing, fold) vectors dealt to the ranks."""
    from treepl_b200 import cv
    flat = build_problem(tips_n)
    tips = np.flatnonzero(flat["child_counts"] == 0)
    grid = [float(v) for v in 10.0 ** np.linspace(3, -1, 16)]
    groups = cv.sample_random_groups(flat, tips, n_groups=10, frac=0.1, seed=1)
    for _ in range(warm + 1):
        res, dt = rk.timed(lambda: cv.cross_validate(flat, grid, groups, randomcv=True, device=rk.local, dist=rk.dist,
                                                     max_iter=max(1, cv_steps // 20), check_every=25))
    n = flat["parents"].shape[0]
    return {"workload": "synthetic %d-tip Yule tree (N=%d): random-subsample CV, 10 folds x 10 %% of the tips, 16 smoothing values "
                        "1e3..1e-1 = 16 full-data fits + 160 fold optimisations, each capped at %d batched evaluations" % (tips_n, n, cv_steps),
            "wall_s": dt, "runs": 1, "n_gpus": rk.world, "chisq": res["chisq"], "best_smoothing": res["best"],
            "full_objective": res["full_objective"], "full_fits_converged": res["full_converged"],
            "fold_vectors_converged": int(rk.vsum(res["cv_converged"])), "fold_vectors": 160,
            "seconds": {"full_fits": res["full_seconds"], "folds": res["cv_seconds"], "scoring_and_gather": res["gather_seconds"]},
            "gpu_launches": int(res["launches"]) * rk.world, "n_local_vectors": res["n_local_vectors"]}


# ---- reference arms of the sharded jobs: BOUNDED SAMPLES of the reference's own optimiser stack on the host cores ----
def ref_sample_restarts(nthreads, budget_starts=None):
    """One optimize_full (utils.cpp:671-763: annealer + TNC rounds, default OptimOptions) per host thread on the first
    `nthreads` of the 1,024 starts; the 1,024-start wall time is extrapolated (starts are independent)."""
    from oracle import refbind as rb
    from treepl_b200 import treeflat as tf
    flat = load_big_test()
    fp, P = tf.generate_param_order_vector(flat["free"], lf=False)
    k = budget_starts or nthreads
    X0 = restart_vectors(flat, P, list(range(k)))
    rp = rb.RefProblem.from_flat(flat)
    rp.set_smoothing(1.0)
    rp.set_freeparams(False)
    f, secs = rp.multistart(X0, nthreads=nthreads)
    rp.close()
    return {"workload": "examples/big_test.cppr8s tree, smoothing 1: %d of the %d random starts, reference optimize_full, one per host thread" % (k, N_STARTS),
            "sample_wall_s": secs, "sample_starts": k, "cores": nthreads,
            "wall_s_extrapolated": secs * N_STARTS / k, "extrapolation": "sample_wall_s x %d / %d (independent starts, all cores busy)" % (N_STARTS, k),
            "objectives": f.tolist(),
            "note": "the reference's optimisers stop far above the optimum on this tree from these starts (b200 arm: 291875.3)"}


def ref_sample_loocv_big(nthreads):
    """Reference CV loop (oracle/ref_harness.cpp::ref_cv = main.cpp:497-822 replayed around the reference's optimize_full) on the
    big_test tree: Langley-Fitch fit + ONE smoothing value's full fit + `nthreads` folds; the 5 x 4,657-fold wall time is extrapolated."""
    from oracle import refbind as rb
    flat = load_big_test()
    tips = [int(t) for t in np.flatnonzero(flat["child_counts"] == 0)]
    rp = rb.RefProblem.from_flat(flat)
    r = rp.cross_validate([10.0], [[t] for t in tips[:nthreads]], nthreads=nthreads)
    rp.close()
    sec = {k: float(v) for k, v in r["seconds"].items()}
    nl, nf = len(BIG_GRID), len(tips)
    extr = sec["lf"] + nl * sec["full"] + nl * sec["folds"] * nf / nthreads
    return {"workload": "LOOCV of the examples/big_test.cppr8s tree: Langley-Fitch fit, 1 of %d smoothing values, %d of %d folds, %d host threads" % (nl, nthreads, nf, nthreads),
            "sample_seconds": sec, "cores": nthreads, "wall_s_extrapolated": extr,
            "extrapolation": "lf + %d x full + %d x folds x %d / %d" % (nl, nl, nf, nthreads),
            "full_objective_sample": r["full_pl"].tolist()}


def ref_loocv_m1155(nthreads):
    from oracle import refbind as rb
    gd, cals, tips = load_m1155()
    p = rb.RefProblem.from_newick(gd.meta["newick"], gd.meta["numsites"], cals, seed=1)
    t = time.perf_counter()
    r = p.cross_validate(M1155_GRID, [[t_] for t_ in tips], nthreads=nthreads)
    dt = time.perf_counter() - t
    p.close()
    return {"workload": "examples/M1155.cppr8s LOOCV (59 folds x 5 smoothing values): the whole job", "wall_s": dt, "cores": nthreads,
            "kind": "reference CV loop replayed around the reference's optimize_full (annealer + TNC / NLopt)",
            "chisq": r["chisq"].tolist(), "full_objective": r["full_pl"].tolist()}


def reference_line(args, metric, value, unit, workload, sample, nthreads, hib=False, scaling="strong", data="synthetic", extra=None):
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * value if unit == "s" else None, "higher_is_better": hib, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": data, "config": {"workload": workload},
            "cpu_baseline": {"value": value, "unit": unit, "cores": nthreads, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line.update(extra or {})
    return line


def init_ranks():
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the PL engine has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return Ranks(dist, local)


def run_job_workload(args):
    """--workload loocv | loocv_big | restarts | restarts_cfg4 | cv100k: the sharded job is the measured workload."""
    rank = int(os.environ.get("RANK", "0"))
    w = args.workload
    metric = {"loocv": "loocv_wall_seconds", "loocv_big": "loocv_wall_seconds", "restarts": "multistart_wall_seconds",
              "restarts_cfg4": "multistart_wall_seconds", "cv100k": "random_cv_wall_seconds"}[w]
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import refbind as rb
        if not rb.available():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libtreepl_ref.so is not built"}))
            return
        nthreads = max(1, os.cpu_count() or 1)
        if w == "loocv":
            r = ref_loocv_m1155(nthreads)
            print(json.dumps(reference_line(args, metric, r["wall_s"], "s", r["workload"], "the whole workload, %d OpenMP threads over folds" % nthreads,
                                            nthreads, data="examples/M1155.tre", extra={"loocv": r})))
        elif w == "loocv_big":
            r = ref_sample_loocv_big(nthreads)
            print(json.dumps(reference_line(args, metric, r["wall_s_extrapolated"], "s", r["workload"], r["extrapolation"] + " (EXTRAPOLATED from a bounded sample)",
                                            nthreads, data="examples/comm_mol_rr.tre", extra={"loocv_big": r})))
        elif w == "restarts":
            r = ref_sample_restarts(nthreads)
            print(json.dumps(reference_line(args, metric, r["wall_s_extrapolated"], "s", r["workload"], r["extrapolation"] + " (EXTRAPOLATED from a bounded sample)",
                                            nthreads, data="examples/comm_mol_rr.tre", extra={"restarts": r})))
        else:
            print(json.dumps({"impl": "reference", "unavailable": "cv100k: the reference needs ~11 min of O(N^2) set-up and ~14 ms per evaluation per thread "
                              "on this tree (SURVEY.md 6); restarts_cfg4: its optimize_full does not move from these starts under log_pen"}))
        return
    rk = init_ranks()
    sampler = ClockSampler(rk.local)
    if rk.rank == 0:
        sampler.start()
    if w == "loocv":
        r = job_loocv(rk, "M1155", runs=args.steps, warm=args.warmup)
        data = "examples/M1155.tre (tests/golden/M1155.npz)"
    elif w == "loocv_big":
        r = job_loocv(rk, "big_test", runs=args.steps, warm=args.warmup)
        data = "examples/comm_mol_rr.tre (tests/golden/big_test.npz)"
    elif w in ("restarts", "restarts_cfg4"):
        r = job_restarts(rk, runs=args.steps, warm=args.warmup, config4=(w == "restarts_cfg4"))
        data = "examples/comm_mol_rr.tre (tests/golden/big_test.npz), synthetic random starts"
    else:
        r = job_cv100k(rk, args.tips, args.cv_steps, warm=args.warmup)
        data = "synthetic"
    clocks = sampler.stop() if rk.rank == 0 else None
    if rk.rank == 0:
        P = r.get("P", 13967 if w.startswith("restarts") else 299996)
        nloc = r.get("n_local_vectors", r.get("n_local", 0))
        line = {"metric": metric, "value": r["wall_s"], "unit": "s", "n_gpus": rk.world, "steps": r["runs"], "warmup": args.warmup,
                "ms_per_step": 1e3 * r["wall_s"], "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": data,
                "config": {"workload": r["workload"], "optimizer": "tpl_optimize_batch method 1 (batched truncated Newton, analytic Hessian-vector kernel, "
                           "block-tree preconditioner, CUDA graph)", "parallelism": "independent optimisations dealt to %d GPU(s); all-gather of chi-square terms / "
                           "best objectives" % rk.world},
                "gpu_launches": r["gpu_launches"], "clocks": clocks,
                "e2e": {"value": r["wall_s"], "unit": "s", "h2d_bytes_per_step": int(8 * P * nloc), "d2h_bytes_per_step": int(8 * P * nloc)}}
        line.update({k: v for k, v in r.items() if k not in ("workload", "wall_s", "runs", "gpu_launches")})
        print(json.dumps(line))
    if rk.dist is not None:
        rk.dist.destroy_process_group()


def run_sweep(args):
    """--workload sweep: (1) evaluations/s and roofline fraction of the batched device-resident
    entry at B in {1, 2, 16, 32, 64, 256, 1024} on the 100k-tip tree; (2) the B = 1 CALLBACK cost the reference's optimisers
    pay (tpl_calc_pl_grad with host buffers: H2D of x, the kernels, D2H of f and g, one call per trial point - src/tnc.c:328)
    on M1155 / big_test / the 100k-tip tree next to the reference's own calc_pl + calc_gradient on one host thread."""
    import torch
    from treepl_b200 import engine as eng
    from treepl_b200 import treeflat as tf
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import Golden
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the PL engine has no CPU fallback")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    peak, peak_src = measured_peaks()
    flat = build_problem(args.tips)
    n = flat["parents"].shape[0]
    fp, P = tf.generate_param_order_vector(flat["free"], lf=False)
    plp = eng.PLCalcParallel.from_flat(flat, device=0)
    plp.set_freeparams(P, False, fp)
    plp.smoothing = 1.0
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    rows = []
    for B in (1, 2, 16, 32, 64, 256, 1024):
        plp.batch_configure(B)
        X = torch.from_numpy(make_vectors(flat, P, B, 1000)).to(dev)
        F = torch.empty(B, dtype=torch.float64, device=dev)
        G = torch.empty((P, B), dtype=torch.float64, device=dev)
        step = lambda: plp.calc_pl_grad_batch_dev(B, X.data_ptr(), F.data_ptr(), G.data_ptr(), eng.GRAD_AD, tstream.cuda_stream)
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        steps = 50 if B >= 64 else 200
        plp.set_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        cnt, main_ms, fin_ms = plp.get_timing()
        plp.set_timing(False)
        bytes_per_launch = (16.0 * P + 25.0 * n / B) * B
        rows.append({"B": B, "ms_per_step": ms, "evals_per_s": B / (ms * 1e-3), "kernel_generation": plp.last_kernel_generation(),
                     "launches_per_step": plp.last_launch_count(), "main_kernel_ms": main_ms / max(cnt, 1),
                     "roofline_frac_step": bytes_per_launch / (ms * 1e-3) / 1e9 / peak,
                     "roofline_frac_main_kernel": bytes_per_launch / (main_ms / max(cnt, 1) * 1e-3) / 1e9 / peak})
        del X, F, G
    plp.close()
    torch.cuda.set_stream(torch.cuda.default_stream(dev))
    # ---- B = 1 callback latency through the host-buffer single-point entry ----
    from oracle import refbind as rb
    cb = []
    for name, fl in (("M1155", dict(Golden("M1155").flat)), ("big_test", dict(Golden("big_test").flat)), ("synthetic %d tips" % args.tips, flat)):
        fpk, Pk = tf.generate_param_order_vector(fl["free"], lf=False)
        q = eng.PLCalcParallel.from_flat(fl, device=0)
        q.set_freeparams(Pk, False, fpk)
        q.smoothing = 1.0
        x = np.ascontiguousarray(make_vectors(fl, Pk, 1, 7)[:, 0]) if name.startswith("synthetic") else None
        if x is None:
            from treepl_b200 import cv
            x = cv.start_vector(fl, Pk, len(fl["parents"]) - 1)
        for _ in range(20):
            q.calc_pl_grad(x, eng.GRAD_ANALYTIC)
        reps = 300
        t = time.perf_counter()
        for _ in range(reps):
            q.calc_pl_grad(x, eng.GRAD_ANALYTIC)
        us = 1e6 * (time.perf_counter() - t) / reps
        # the same callback as the reference's function_plcp issues it through the C++ mirror: calc_pl, then calc_gradient
        # (optimize_tnc.cpp:83-86) - one evaluation, the gradient rides along with calc_pl once the pair pattern is seen
        for _ in range(3):
            q.calc_pl(x); q.calc_gradient(x)
        t = time.perf_counter()
        for _ in range(reps):
            q.calc_pl(x); q.calc_gradient(x)
        pair_us = 1e6 * (time.perf_counter() - t) / reps
        row = {"tree": name, "N": int(len(fl["parents"])), "P": int(Pk), "b200_callback_us": us, "b200_calc_pl_then_calc_gradient_us": pair_us,
               "kernel_generation": q.last_kernel_generation(), "launches_per_call": q.last_launch_count()}
        q.close()
        if rb.available():
            rp = rb.RefProblem.from_flat(fl)
            rp.set_smoothing(1.0)
            rp.set_freeparams(False)
            t1 = rp.time_evals(x, 1, 1, True)
            r = int(max(3, min(20000, 1.0 / max(t1, 1e-6))))
            row["reference_one_thread_us"] = 1e6 * rp.time_evals(x, r, 1, True) / r
            rp.close()
        cb.append(row)
    print(json.dumps({"metric": "pl_obj_grad_evals_per_sec_vs_B", "unit": "evals/s", "n_gpus": 1, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": "B sweep on the synthetic %d-tip tree (N=%d, P=%d), AD gradient, operands resident in HBM; "
                                             "B=1 callback latency through tpl_calc_pl_grad (host buffers)" % (args.tips, n, P)},
                      "peak_gbs": peak, "peak_source": peak_src, "sweep": rows, "callback": cb}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--tips", type=int, default=100000)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--mode", default="ad", choices=["ad", "analytic"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--min-warm-seconds", type=float, default=1.0)
    ap.add_argument("--workload", default="evals", choices=["evals", "loocv", "loocv_big", "restarts", "restarts_cfg4", "cv100k", "sweep"])
    ap.add_argument("--cv-steps", type=int, default=8000, help="cv100k: cap on batched evaluations per optimisation")
    ap.add_argument("--no-sharded", "--no-loocv", dest="no_sharded", action="store_true",
                    help="skip the sharded optimiser-level jobs attached to the default line")
    args = ap.parse_args()
    if args.workload == "sweep":
        run_sweep(args)
        return
    if args.workload != "evals":
        big = args.workload != "loocv"
        if args.steps == 200:
            args.steps = 1 if (big or args.impl == "reference") else 5
        if args.warmup == 3:
            args.warmup = 0 if args.impl == "reference" else (1 if big else 3)
        run_job_workload(args)
        return
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    flat = build_problem(args.tips)
    n = flat["parents"].shape[0]
    from treepl_b200 import treeflat as tf
    fp, P = tf.generate_param_order_vector(flat["free"], lf=False)

    if args.impl == "reference":
        run_reference_arm(args, flat, P)
        return

    import torch
    from treepl_b200 import engine as eng

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the PL engine has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    B = args.batch
    mode = eng.GRAD_AD if args.mode == "ad" else eng.GRAD_ANALYTIC
    plp = eng.PLCalcParallel.from_flat(flat, device=local)
    plp.set_freeparams(P, False, fp)
    plp.smoothing = 1.0
    plp.batch_configure(B)
    Xh = make_vectors(flat, P, B, 1000 + rank)
    dev = torch.device("cuda", local)
    X = torch.from_numpy(Xh).to(dev)
    F = torch.empty(B, dtype=torch.float64, device=dev)
    G = torch.empty((P, B), dtype=torch.float64, device=dev)
    # a non-default torch stream: its handle is what the C ABI launches on, and what the events time
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step():
        plp.calc_pl_grad_batch_dev(B, X.data_ptr(), F.data_ptr(), G.data_ptr(), mode, stream)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # warm-up: at least W steps and at least ~1 s of load, so that the clock sampler (100 ms period)
    # sees the GPU under this kernel's load before and during the timed region
    t_w = time.perf_counter()
    done = 0
    while done < args.warmup or (time.perf_counter() - t_w) < args.min_warm_seconds:
        step()
        done += 1
        if done % 16 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    # correctness guard inside the bench: finite objective for every vector
    Fh = F.cpu().numpy()
    if not (np.isfinite(Fh).all() and (Fh < 1e15).all()):
        raise SystemExit("bench.py: engine returned infeasible/NaN objectives on the bench vectors")

    # checker leg (outside the timed region, rank 0): two of the bench's own vectors against the CPU oracle
    parity = None
    if rank == 0:
        parity = oracle_parity(flat, Xh, Fh, G, mode == eng.GRAD_AD, (0, B - 1))
        if parity["f_rel"] > 1e-10 or parity["g_rel"] > 1e-8:
            raise SystemExit("bench.py: the bench vectors fail the parity gate against the oracle: %r" % (parity,))

    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    # per-kernel durations for the roofline: the same K steps again, right behind the timed region, with CUDA events recorded
    # inside the library around the main kernel (kept out of the timed region: two more event records per step)
    plp.set_timing(True)
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    cnt, main_ms, fin_ms = plp.get_timing()
    plp.set_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = world * args.steps * plp.last_launch_count()

    # ---- end to end through the C ABI with pinned host buffers ----
    e2e_steps = max(1, args.e2e_steps)
    Xp = torch.from_numpy(Xh).pin_memory()
    Fp = torch.empty(B, dtype=torch.float64).pin_memory()
    Gp = torch.empty((P, B), dtype=torch.float64).pin_memory()
    xp, fpn, gp = Xp.numpy(), Fp.numpy(), Gp.numpy()
    L = plp._L
    import ctypes as C
    dptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for _ in range(2):                                                           # warm: staging buffers, streams, first-touch of the pinned pages
        L.tpl_calc_pl_grad_batch(plp._h, B, dptr(xp), dptr(fpn), dptr(gp), mode)
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rc = L.tpl_calc_pl_grad_batch(plp._h, B, dptr(xp), dptr(fpn), dptr(gp), mode)
        if rc < 0:
            raise SystemExit("e2e call failed: " + L.tpl_last_error().decode())
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    # the host path evaluates 64-vector sub-batches (pipelined with the copies): the objective partials are
    # grouped differently, so agreement is to rounding, not bitwise
    if not np.allclose(fpn, Fh, rtol=1e-13, atol=0):
        raise SystemExit("bench.py: host-buffer path and device-buffer path disagree")

    if rank == 0:
        peak, peak_src = measured_peaks()
        evals = B * args.steps
        value = world * evals / (ms * 1e-3)
        bytes_per_eval = 16.0 * P + 25.0 * n / B
        main_s = main_ms * 1e-3 / max(cnt, 1)
        achieved = bytes_per_eval * B / main_s / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "synthetic %d-tip Yule tree, N=%d nodes, P=%d params, PL (lf=false), smoothing=1, "
                                       "B=%d vectors per step per GPU, gradient=%s; X+G = %.0f MB per step (> 126 MB L2, "
                                       "no flush needed)" % (args.tips, n, P, B, args.mode, 2 * P * B * 8 / 1e6),
                           "batch_per_gpu": B, "parallelism": "independent vectors sharded over %d GPU(s)" % world},
                "gpu_launches": launches,
                "parity": parity,
                "clocks": clocks,
                "e2e": {"value": world * B * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(P * B * 8),
                        "d2h_bytes_per_step": int(P * B * 8 + B * 8), "steps": e2e_steps},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": measured_traffic(args.tips, B, args.mode), "peak_source": peak_src,
                             "kernel": {4: "pl_desc_kernel", 3: "pl_persist_kernel", 2: "pl_tile_kernel",
                                        1: "pl_main_kernel"}.get(plp.last_kernel_generation(), "?"),
                             "kernel_ms": main_ms / max(cnt, 1), "finalize_ms": fin_ms / max(cnt, 1),
                             "algorithmic_bytes_per_launch": bytes_per_eval * B,
                             "frac_of_whole_step": bytes_per_eval * B / (ms / args.steps * 1e-3) / 1e9 / peak,
                             "how": "kernel_ms: CUDA events around the main kernel, %d launches right behind the timed region; "
                                    "frac_of_whole_step: the timed region itself (main + finalize kernels per step)" % max(cnt, 1)}}
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = cpu_reference_rate(flat, np.ascontiguousarray(Xh[:, 0]))
                if r is not None:
                    line["cpu_baseline"] = {"value": r[0], "unit": UNIT, "cores": r[1], "kind": "reference", "sample": r[2]}
            except Exception as ex:      # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"error": repr(ex)}
    plp.close()
    del X, G, F, Xp, Gp, Fp
    torch.cuda.empty_cache()
    torch.cuda.set_stream(torch.cuda.default_stream(dev))
    # ---- the workloads north_star shards over the GPUs, at EVERY N: optimiser-level jobs through the device-resident
    # batched optimiser (strong scaling: the same job on N GPUs), wall seconds incl. all host work of the job ----
    if not args.no_sharded:
        rk = Ranks(dist, local)
        sh = {}
        # each job once untimed (first use of its work space: tens of GB of device memory are mapped), then once timed
        for name, fn in (("loocv_M1155", lambda: job_loocv(rk, "M1155", runs=1, warm=1)),
                         ("restarts", lambda: job_restarts(rk, runs=1, warm=1)),
                         ("loocv_big_test", lambda: job_loocv(rk, "big_test", runs=1, warm=1)),
                         ("cv100k", lambda: job_cv100k(rk, args.tips, args.cv_steps, warm=1))):
            try:
                sh[name] = fn()
            except Exception as ex:                  # collective error agreement happens inside cv.*: every rank raises
                sh[name] = {"error": repr(ex)}
        if rank == 0:
            line["sharded"] = sh
            line["loocv"] = sh["loocv_M1155"]
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
