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

Second workload (BASELINE.json: "LOO-CV wall time at 1/8 B200"), not the driver's default:
    python bench.py --workload loocv [--gpus N] [--steps K] [--impl reference]
one step = the complete leave-one-out cross-validation of examples/M1155.cppr8s (59 folds x 5 smoothing values,
tests/golden/M1155.npz) with the library's device-resident batched L-BFGS; the (smoothing, fold) vectors are
dealt to the ranks and the chi-square terms all-gathered over NCCL (strong scaling).  The reference arm replays
the reference's own CV loop around its optimize_full (annealer + TNC) on the host cores (oracle/_ref).
    python bench.py --workload restarts [--gpus N]
one step = 1,024 random-restart optimisations of the examples/big_test.cppr8s tree (BASELINE.json config 4), starts
dealt to the ranks, best-of-restarts exchange over NCCL.
    python bench.py --workload cv100k [--gpus N] [--cv-steps S]
random-subsample CV over 16 smoothing values on the synthetic 100k-tip tree (BASELINE.json config 5), capped.
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
    try:
        np.savez(cache, **d)
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
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
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
    if not args.no_loocv:
        # second half of BASELINE.json's metric: LOO-CV wall time (one run of the reference's CV loop, all host threads)
        try:
            gd, cals, tips = load_m1155()
            p = rb.RefProblem.from_newick(gd.meta["newick"], gd.meta["numsites"], cals, seed=1)
            t = time.perf_counter()
            r = p.cross_validate(M1155_GRID, [[t_] for t_ in tips], nthreads=nthreads)
            line["loocv"] = {"workload": "examples/M1155.cppr8s LOOCV (59 folds x 5 smoothing values)",
                             "wall_s": time.perf_counter() - t, "cores": nthreads,
                             "kind": "reference CV loop replayed around the reference's optimize_full (annealer + TNC)",
                             "full_objective": r["full_pl"].tolist()}
            p.close()
        except Exception as ex:
            line["loocv"] = {"error": repr(ex)}
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


def run_loocv(args):
    """--workload loocv: wall time of the complete M1155 leave-one-out cross-validation."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    gd, cals, tips = load_m1155()
    workload = "examples/M1155.cppr8s: LOOCV, %d folds x %d smoothing values = %d optimisations of P=170 parameters" % (
        len(tips), len(M1155_GRID), len(tips) * len(M1155_GRID))
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import refbind as rb
        if not rb.available():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libtreepl_ref.so is not built"}))
            return
        nthreads = max(1, os.cpu_count() or 1)
        times, last = [], None
        for s in range(args.warmup + args.steps):
            p = rb.RefProblem.from_newick(gd.meta["newick"], gd.meta["numsites"], cals, seed=1)
            t = time.perf_counter()
            last = p.cross_validate(M1155_GRID, [[t_] for t_ in tips], nthreads=nthreads)
            dt = time.perf_counter() - t
            p.close()
            if s >= args.warmup:
                times.append(dt)
        val = sum(times) / len(times)
        sample = ("the whole workload: reference Langley-Fitch fit + CV loop (main.cpp:497-822 replayed) around the "
                  "reference's optimize_full (annealer + TNC; NLopt not built), %d OpenMP threads over folds" % nthreads)
        print(json.dumps({"impl": "reference", "metric": "loocv_wall_seconds", "value": val, "unit": "s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * val, "higher_is_better": False,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "examples/M1155.tre",
                          "config": {"workload": workload},
                          "chisq": last["chisq"].tolist(), "full_objective": last["full_pl"].tolist(),
                          "note": "TNC-only replay: its optimiser stops short of the optimum (compare full_objective with the "
                                  "b200 arm's), so its chi-square table is NOT the parity reference; SURVEY.md 8c's is",
                          "cpu_baseline": {"value": val, "unit": "s", "cores": nthreads, "kind": "reference", "sample": sample},
                          "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    import torch
    from treepl_b200 import cv
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the PL engine has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    times, res = [], None
    for s in range(args.warmup + args.steps):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        res = cv.loocv(dict(gd.flat), M1155_GRID, tips, device=local, max_iter=6000, dist=dist)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        dt = time.perf_counter() - t
        if dist is not None:
            tt = torch.tensor([dt], dtype=torch.float64, device=torch.device("cuda", local))
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        if s >= args.warmup:
            times.append(dt)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        val = sum(times) / len(times)
        chisq = res["chisq"]
        rel = [abs(a - b) / b for a, b in zip(chisq, M1155_REF_CHISQ)]
        print(json.dumps({"metric": "loocv_wall_seconds", "value": val, "unit": "s", "n_gpus": world, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * val, "higher_is_better": False, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "examples/M1155.tre (tests/golden/M1155.npz)",
                          "config": {"workload": workload, "optimizer": "tpl_optimize_batch method 1 (batched truncated Newton, analytic Hessian-vector kernel, CUDA graph)",
                                     "parallelism": "(smoothing, fold) vectors dealt to %d GPU(s), chi-square all-gather" % world},
                          "gpu_launches": int(res["launches"]) * world, "clocks": clocks,
                          "chisq": chisq, "chisq_reference_run": M1155_REF_CHISQ, "chisq_max_rel_diff": max(rel),
                          "best_smoothing": res["best"], "full_objective": res["full_objective"],
                          "device_ms": {"full_fits": res["full_device_ms"], "folds": res["cv_device_ms"]},
                          "batched_evaluations": {"full_fits": res["full_info"]["nfev"], "folds": res["cv_nfev"]},
                          "hessian_vector_products": res["cg_iterations"],
                          "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": int(2 * 170 * 8 * (6 + res["n_local_vectors"])),
                                  "d2h_bytes_per_step": int(2 * 170 * 8 * (6 + res["n_local_vectors"]))}}))
    if dist is not None:
        dist.destroy_process_group()


def run_restarts(args):
    """--workload restarts: BASELINE.json config 4 - 1,024 batched random-restart optimisations of examples/big_test.cppr8s
    (N = 9,313 nodes, P = 13,967 parameters) sharded over the GPUs, best-of-restarts exchange over NCCL."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import Golden
    flat = dict(Golden("big_test").flat)
    n_starts = 1024
    workload = "examples/big_test.cppr8s tree (N=9313, P=13967), smoothing 1: %d random-restart optimisations to |g| <= 1e-8 |f|" % n_starts
    if args.impl == "reference":
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "the reference has no batched multi-start driver; one big_test fit "
                              "with its optimize_full takes minutes per start on one core (SURVEY.md 6)"}))
        return
    import torch
    from treepl_b200 import cv
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the PL engine has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    times, res = [], None
    for s in range(args.warmup + args.steps):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        res = cv.multistart(flat, 1.0, n_starts, device=local, dist=dist, check_every=50)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        dt = time.perf_counter() - t
        if dist is not None:
            tt = torch.tensor([dt], dtype=torch.float64, device=torch.device("cuda", local))
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        if s >= args.warmup:
            times.append(dt)
    clocks = sampler.stop() if rank == 0 else None
    conv = res["local_converged"]
    if dist is not None:
        tt = torch.tensor([conv], dtype=torch.int64, device=torch.device("cuda", local))
        dist.all_reduce(tt)
        conv = int(tt.item())
    if rank == 0:
        val = sum(times) / len(times)
        print(json.dumps({"metric": "multistart_wall_seconds", "value": val, "unit": "s", "n_gpus": world, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * val, "higher_is_better": False, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "examples/big_test.tre (tests/golden/big_test.npz), synthetic random starts",
                          "config": {"workload": workload, "optimizer": "tpl_optimize_batch method 1 (batched truncated Newton)",
                                     "parallelism": "starts dealt to %d GPU(s); all-gather of best objectives + broadcast of the winner" % world},
                          "optimisations_per_second": n_starts / val, "converged": conv, "best_objective": res["best"],
                          "winner_rank": res["winner_rank"], "gpu_launches": int(res["launches"]) * world, "clocks": clocks,
                          "batched_evaluations": res["nfev"], "hessian_vector_products_rank0": res["cg_iterations"],
                          "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": int(13967 * 8 * res["n_local"]),
                                  "d2h_bytes_per_step": int(13967 * 8 * res["n_local"])}}))
    if dist is not None:
        dist.destroy_process_group()


def run_cv100k(args):
    """--workload cv100k: BASELINE.json config 5 - random-subsample cross-validation (10 folds of 10 % of the tips,
    src/main.cpp:609-656) over 16 smoothing values 10^3 ... 10^-1 on the synthetic 100,000-tip tree, (smoothing, fold)
    vectors dealt to the GPUs.  Every optimisation is capped at `--cv-steps` batched evaluations (x 4 CG iterations):
    with the Jacobi preconditioner the vectors with smoothing >= ~50 are NOT converged at the default cap and the line
    says how many are."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "the reference needs ~11 min of O(N^2) set-up and ~14 ms per "
                              "evaluation per thread on this tree (SURVEY.md 6): its CV over 176 optimisations is not runnable"}))
        return
    import torch
    from treepl_b200 import cv
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the PL engine has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    flat = build_problem(args.tips)
    tips = np.flatnonzero(flat["child_counts"] == 0)
    grid = [float(v) for v in 10.0 ** np.linspace(3, -1, 16)]
    groups = cv.sample_random_groups(flat, tips, n_groups=10, frac=0.1, seed=1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    res = cv.cross_validate(flat, grid, groups, randomcv=True, device=local, dist=dist, max_iter=max(1, args.cv_steps // 20), check_every=25)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    dt = time.perf_counter() - t
    conv = res["cv_converged"]
    if dist is not None:
        tt = torch.tensor([dt, float(conv)], dtype=torch.float64, device=torch.device("cuda", local))
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt)
        dt, conv = float(mx[0].item()), int(tt[1].item())
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        n = flat["parents"].shape[0]
        print(json.dumps({"metric": "random_cv_wall_seconds", "value": dt, "unit": "s", "n_gpus": world, "steps": 1, "warmup": 0,
                          "ms_per_step": 1e3 * dt, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic",
                          "config": {"workload": "synthetic %d-tip Yule tree (N=%d): random-subsample CV, 10 folds x 10 %% of the tips, "
                                                 "16 smoothing values 1e3..1e-1 = 16 full-data fits + 160 fold optimisations, each capped at "
                                                 "%d batched evaluations" % (args.tips, n, args.cv_steps),
                                     "optimizer": "tpl_optimize_batch method 1 (batched truncated Newton)",
                                     "parallelism": "(smoothing, fold) vectors dealt to %d GPU(s), chi-square all-gather" % world},
                          "chisq": res["chisq"], "best_smoothing": res["best"], "full_objective": res["full_objective"],
                          "full_fits_converged": res["full_converged"], "fold_vectors_converged": conv,
                          "seconds": {"full_fits": res["full_seconds"], "folds": res["cv_seconds"], "scoring_and_gather": res["gather_seconds"]},
                          "gpu_launches": int(res["launches"]) * world, "clocks": clocks,
                          "e2e": {"value": dt, "unit": "s", "h2d_bytes_per_step": int(8 * 299996 * (16 + res["n_local_vectors"])),
                                  "d2h_bytes_per_step": int(8 * 299996 * (16 + res["n_local_vectors"]))}}))
    if dist is not None:
        dist.destroy_process_group()


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
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--min-warm-seconds", type=float, default=1.0)
    ap.add_argument("--workload", default="evals", choices=["evals", "loocv", "restarts", "cv100k"])
    ap.add_argument("--cv-steps", type=int, default=8000, help="cv100k: cap on batched evaluations per optimisation")
    ap.add_argument("--no-loocv", action="store_true", help="skip the LOO-CV summary attached to the default line")
    args = ap.parse_args()
    if args.workload == "cv100k":
        run_cv100k(args)
        return
    if args.workload == "restarts":
        if args.steps == 200:
            args.steps = 2
        if args.warmup == 3:
            args.warmup = 1
        run_restarts(args)
        return
    if args.workload == "loocv":
        if args.steps == 200:
            args.steps = 1 if args.impl == "reference" else 5
        if args.impl == "reference" and args.warmup == 3:
            args.warmup = 0
        run_loocv(args)
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

    plp.set_timing(True)
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
    L.tpl_calc_pl_grad_batch(plp._h, B, dptr(xp), dptr(fpn), dptr(gp), mode)     # warm (allocations)
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
                             "algorithmic_bytes_per_launch": bytes_per_eval * B}}
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = cpu_reference_rate(flat, np.ascontiguousarray(Xh[:, 0]))
                if r is not None:
                    line["cpu_baseline"] = {"value": r[0], "unit": UNIT, "cores": r[1], "kind": "reference", "sample": r[2]}
            except Exception as ex:      # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"error": repr(ex)}
        if world == 1 and not args.no_loocv:
            try:
                from treepl_b200 import cv
                gd, cals, tips = load_m1155()
                cv.loocv(dict(gd.flat), M1155_GRID, tips, device=local, max_iter=6000)        # warm (allocations, graphs)
                t = time.perf_counter()
                r = cv.loocv(dict(gd.flat), M1155_GRID, tips, device=local, max_iter=6000)
                wall = time.perf_counter() - t
                line["loocv"] = {"workload": "examples/M1155.cppr8s LOOCV (59 folds x 5 smoothing values), device-resident batched truncated Newton",
                                 "wall_s": wall, "chisq": r["chisq"], "chisq_reference_run": M1155_REF_CHISQ,
                                 "chisq_max_rel_diff": max(abs(a - b) / b for a, b in zip(r["chisq"], M1155_REF_CHISQ)),
                                 "best_smoothing": r["best"], "full_objective": r["full_objective"],
                                 "gpu_launches": int(r["launches"]), "batched_evaluations": r["full_info"]["nfev"] + r["cv_nfev"]}
            except Exception as ex:
                line["loocv"] = {"error": repr(ex)}
        print(json.dumps(line))
    plp.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
