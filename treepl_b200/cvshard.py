"""Sharding of independent PL work items over the GPUs of one box.

Work item = one (smoothing value, CV fold) pair or one random restart: the reference runs the folds in an
OpenMP `parallel for ... reduction(+:chisq)` (src/main.cpp:689) and the restarts one after the other
(src/utils.cpp:475-553).  They are independent, so each rank (one process per GPU) evaluates its own items
as the batch dimension of the PL kernels; the tree (~25 N bytes) is replicated.  The only exchanges are
  * per smoothing value: all-gather of the per-fold chi-square terms (8 bytes x folds), summed in FOLD ORDER on
    every rank (deterministic, unlike the OpenMP reduction), and
  * for restarts: all-gather of (best objective, rank), then a broadcast of the winner's parameter vector.
Both are latency-bound collectives with nothing to fuse into a kernel; NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_items(n_items, rank, world):
    """Static block-cyclic assignment: item i belongs to rank i % world (folds of similar cost interleave)."""
    return list(range(rank, n_items, world))


def owner_of(item, world):
    return item % world


def gather_fold_scores(local_scores, n_items, dist=None, device=None):
    """local_scores: {item: chi-square term}, or an array of this rank's terms in shard_items order.  Returns the
    full vector (item order) on every rank and its fixed-order sum.  With dist=None (single process) no collective
    is issued."""
    import torch

    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    per = (n_items + world - 1) // world
    if isinstance(local_scores, dict):
        vals = np.array([float(local_scores[item]) for item in shard_items(n_items, rank, world)], dtype=np.float64)
    else:
        vals = np.asarray(local_scores, dtype=np.float64)
    host = np.full(per, np.nan)
    host[: vals.shape[0]] = vals
    if dist is None:
        parts = [host]
    else:
        mine = torch.as_tensor(host, device=device)
        got = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(got, mine)
        parts = [t.cpu().numpy() for t in got]
    full = np.full(n_items, np.nan)
    for r, part in enumerate(parts):
        cnt = len(range(r, n_items, world))
        full[r::world] = part[:cnt]
    total = 0.0
    for v in full:                    # fixed item order -> identical on every rank and for every world size
        total += v
    return full, total


def gather_vectors(local_vectors, n_items, length, dist=None, device=None):
    """local_vectors: {item: 1-D array of `length` doubles} for this rank's items (shard_items order).
    Returns the [length, n_items] matrix of all items on every rank (all-gather of length x per-rank-items doubles:
    the full-data optima every fold of a smoothing value starts from)."""
    import torch

    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    per = (n_items + world - 1) // world
    host = np.zeros((per, length))
    for k, item in enumerate(shard_items(n_items, rank, world)):
        host[k] = np.asarray(local_vectors[item], dtype=np.float64)
    if dist is None:
        parts = [host]
    else:
        mine = torch.as_tensor(host, device=device)
        got = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(got, mine)
        parts = [t.cpu().numpy() for t in got]
    out = np.zeros((length, n_items))
    for r, part in enumerate(parts):
        cnt = len(range(r, n_items, world))
        out[:, r::world] = part[:cnt].T
    return out


def best_of_restarts(best_f, best_x, dist=None, device=None):
    """Each rank passes its best (objective, parameter vector); returns the global best on every rank.
    Ties go to the lowest rank."""
    import torch

    if dist is None:
        return float(best_f), np.asarray(best_x, dtype=np.float64), 0
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.tensor([float(best_f)], dtype=torch.float64, device=device)
    allf = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allf, mine)
    fs = np.array([t.item() for t in allf])
    fs_cmp = np.where(np.isnan(fs), np.inf, fs)
    winner = int(np.argmin(fs_cmp))
    x = torch.as_tensor(np.asarray(best_x, dtype=np.float64), device=device).clone()
    dist.broadcast(x, src=winner)
    return float(fs[winner]), x.cpu().numpy(), winner


def raise_together(error, dist=None, device=None):
    """Collective error agreement: every rank passes its local exception (or None); if ANY rank failed, every rank
    raises, so no rank is left blocked in the next all-gather / broadcast."""
    if dist is None:
        if error is not None:
            raise error
        return
    import torch

    flag = torch.tensor([0.0 if error is None else 1.0], dtype=torch.float64, device=device)
    flags = [torch.empty_like(flag) for _ in range(dist.get_world_size())]
    dist.all_gather(flags, flag)
    failed = [r for r, t in enumerate(flags) if t.item() != 0.0]
    if error is not None:
        raise error
    if failed:
        raise RuntimeError("optimisation failed on rank(s) %s" % failed)
