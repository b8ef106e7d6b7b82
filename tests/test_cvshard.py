"""N>1 path of the sharded CV / restart workloads on CPU: world_size-2 gloo processes."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from treepl_b200 import cvshard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.RandomState(99)
        truth = rng.rand(n_items) * 100            # the chi-square term every fold "computes"
        mine = {i: truth[i] for i in cvshard.shard_items(n_items, rank, world)}
        full, total = cvshard.gather_fold_scores(mine, n_items, dist=dist)
        # restarts: rank r's best objective / vector
        f, x, winner = cvshard.best_of_restarts(10.0 - rank, np.full(5, float(rank)), dist=dist)
        vecs = {i: np.arange(7, dtype=np.float64) + 100 * i for i in cvshard.shard_items(n_items, rank, world)}
        mat = cvshard.gather_vectors(vecs, n_items, 7, dist=dist)
        assert np.array_equal(mat, np.arange(7)[:, None] + 100.0 * np.arange(n_items)[None, :])
        out.put((rank, full.tolist(), total, f, x.tolist(), winner))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [59, 8, 3])
def test_two_rank_gloo_gather_matches_single_process(n_items):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.RandomState(99)
    truth = rng.rand(n_items) * 100
    full1, total1 = cvshard.gather_fold_scores({i: truth[i] for i in range(n_items)}, n_items)
    for rank, full, total, f, x, winner in res:
        assert np.array_equal(np.array(full), truth)
        assert total == total1                      # bitwise: fixed-order sum, independent of world size
        assert winner == 1 and f == 9.0 and x == [1.0] * 5


def test_shard_items_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in cvshard.shard_items(59, r, world))
        assert seen == list(range(59))
        sizes = [len(cvshard.shard_items(59, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def _raise_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        try:
            cvshard.raise_together(ValueError("start point infeasible") if rank == 1 else None, dist=dist)
            out.put((rank, "no error"))
        except ValueError as ex:
            out.put((rank, "ValueError"))
        except RuntimeError as ex:
            out.put((rank, str(ex)))
        cvshard.raise_together(None, dist=dist)            # and the group is still usable afterwards
    finally:
        dist.destroy_process_group()


def test_a_failure_on_one_rank_is_raised_on_every_rank():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_raise_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(out.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[1] == "ValueError" and "rank(s) [1]" in res[0]
