"""Multi-GPU host logic on CPU: scenario sharding and the end-of-run statistics gather (gloo, world_size 2)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rnmpc_b200  # noqa: F401
from rnmpc_b200.sharding import gather_statistics, shard_bounds


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 4096, 65536, 65537):
        for w in (1, 2, 3, 4, 8):
            chunks = [shard_bounds(n, r, w) for r in range(w)]
            assert chunks[0][0] == 0 and chunks[-1][1] == n
            assert all(chunks[i][1] == chunks[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in chunks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(1001, rank, world)
    local = dict(solves=(hi - lo) * 29, failed_solves=rank, violations=0, cost_sum=float(np.arange(lo, hi).sum()),
                 max_backoff=0.1 * (rank + 1), qp_iters=10 * (hi - lo), elapsed_ms=5.0 + rank)
    total, per_rank = gather_statistics(local)
    if rank == 0:
        q.put((total, per_rank))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_statistics_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, per_rank = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert total["solves"] == 1001 * 29 and total["failed_solves"] == 1
    assert total["cost_sum"] == float(np.arange(1001).sum())
    assert abs(total["max_backoff"] - 0.2) < 1e-15 and total["elapsed_ms"] == 6.0
    assert len(per_rank) == 2 and per_rank[0]["solves"] + per_rank[1]["solves"] == total["solves"]


def test_gather_statistics_single_process():
    total, per_rank = gather_statistics(dict(solves=5, max_backoff=0.3))
    assert total["solves"] == 5 and total["max_backoff"] == 0.3 and len(per_rank) == 1
