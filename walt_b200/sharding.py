"""Read sharding across GPUs/ranks: contiguous ranges, no data-path collective;
torch.distributed is only used for barriers and the max-over-ranks of timings."""
from __future__ import annotations


def shard_range(n, world, rank):
    """[lo, hi) of rank's contiguous share of n reads: [floor(i*n/G), floor((i+1)*n/G))."""
    return (n * rank) // world, (n * (rank + 1)) // world


def max_over_ranks(values, dist=None, device=None):
    """element-wise MAX of a list of floats over all ranks (identity without a process group)"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(v) for v in values]
    import torch
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def gather_in_order(local, n_total, dist, dtype):
    """Host-side ordered merge of per-rank result slices (rank r holds shard_range(n, G, r))."""
    import numpy as np
    import torch
    world = dist.get_world_size()
    parts = [None] * world
    dist.all_gather_object(parts, local.tobytes())
    out = np.frombuffer(b"".join(parts), dtype=dtype)
    assert out.size == n_total
    return out
