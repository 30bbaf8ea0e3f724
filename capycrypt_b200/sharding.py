"""Host-side partitioning of a batch over ranks / devices: contiguous index ranges balanced by
work, no exchange step.  Pure host logic (numpy); mirrors `split_items` in csrc/hostbatch.h, which the C entry
points use when one ctx spans several GPUs."""
from __future__ import annotations

import numpy as np


def item_cost(off: np.ndarray, rate_bytes: int, per_item: int = 1) -> np.ndarray:
    """Work units per item: permutations for a sponge with the given rate (+1 for the suffix/pad block)."""
    lens = np.diff(np.asarray(off, dtype=np.uint64)).astype(np.int64)
    return lens // rate_bytes + per_item


def contiguous_shards(cost: np.ndarray, world: int) -> list[tuple[int, int]]:
    """[(i0, i1)] per rank: contiguous, covering [0, n), cumulative cost split as evenly as possible."""
    n = len(cost)
    if world <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (max(world, 1) - 1)
    cum = np.concatenate([[0], np.cumsum(cost, dtype=np.int64)])
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r // world
        i = int(np.searchsorted(cum, target, side="left"))
        bounds.append(min(max(i, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def strided_shard(n: int, rank: int, world: int) -> np.ndarray:
    """Indices rank, rank + world, ... : statistically balanced share of a batch with i.i.d. lengths."""
    return np.arange(rank, n, world)


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Timing rule: every multi-rank number is the MAX over ranks."""
    if dist is None:
        return value
    import torch

    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
