"""Multi-GPU partitioning of the ADCensus path: independent stereo pairs are sharded over
ranks (one process per GPU).  There is NO data-path collective -- the vertical scanline
paths, the +-132-row aggregation halo and the raster-order voting leak forbid splitting a
pair -- so torch.distributed is only used to reduce timings and, when asked,
to gather the finished disparity maps on rank 0."""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np


def frames_for_rank(n_frames: int, world: int, rank: int) -> list[int]:
    """Frame i is owned by rank i % world (round-robin keeps ragged batches balanced)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return [i for i in range(n_frames) if i % world == rank]


def max_over_ranks(value: float, dist=None, device=None) -> float:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, dist=None, device=None) -> float:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def run_sharded(pairs: Sequence, compute: Callable, dist=None, gather: bool = True):
    """Runs compute(left, right) on this rank's share of `pairs`; with gather=True rank 0
    returns the full list in frame order (others return None), else every rank returns
    {frame_index: disparity} for its own frames."""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    mine = {i: np.asarray(compute(*pairs[i])) for i in frames_for_rank(len(pairs), world, rank)}
    if not gather:
        return mine
    if world == 1:
        return [mine[i] for i in range(len(pairs))]
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, dst=0)
    if rank != 0:
        return None
    merged: dict = {}
    for p in parts:
        merged.update(p)
    return [merged[i] for i in range(len(pairs))]
