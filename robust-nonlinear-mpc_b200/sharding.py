"""Scenario sharding across the GPUs of one box: independent closed loops, contiguous chunks per rank,
no collective on the solve path, one small statistics gather at the end (torch.distributed: NCCL on GPUs, gloo in the
CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_bounds(n_items: int, rank: int, world_size: int):
    """Contiguous [lo, hi) chunk of `n_items` scenarios owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank outside world")
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


STAT_KEYS = ("solves", "failed_solves", "violations", "cost_sum", "max_backoff", "qp_iters", "elapsed_ms")
_REDUCE = dict(solves="sum", failed_solves="sum", violations="sum", cost_sum="sum", max_backoff="max", qp_iters="sum",
               elapsed_ms="max")


def gather_statistics(local: dict, device=None, group=None):
    """All-gather the per-rank statistics record (a few floats) and reduce it; returns (total, per_rank list).

    Works without an initialised process group (single process) -- then it just wraps `local`."""
    import torch
    import torch.distributed as dist

    vec = torch.tensor([float(local.get(k, 0.0)) for k in STAT_KEYS], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        out = [torch.empty_like(vec) for _ in range(dist.get_world_size(group))]
        dist.all_gather(out, vec, group=group)
        rows = np.stack([o.cpu().numpy() for o in out])
    else:
        rows = vec.cpu().numpy()[None]
    total = {}
    for i, k in enumerate(STAT_KEYS):
        total[k] = float(rows[:, i].sum() if _REDUCE[k] == "sum" else rows[:, i].max())
    per_rank = [dict(zip(STAT_KEYS, map(float, r))) for r in rows]
    return total, per_rank
