"""Multi-GPU plumbing: one process per GPU, torch.distributed for the collectives.

The path shards by film tiles: the reference's spiral tile list is interleaved over the ranks
(tile i -> rank i mod G, the scheme render_manager.rs:206-210 sketches), every rank renders its share into a zeroed
full-size film, and one sum-reduce to rank 0 assembles the image. Tiles are disjoint, so the sum adds each pixel to
zeros only: it is a gather, and the result is bit-identical to the single-rank film. No other data-path collective.
"""
from __future__ import annotations

import numpy as np


def partition_tiles(tiles: np.ndarray, rank: int, world: int) -> np.ndarray:
    """This rank's share of the spiral tile list, centre-out order preserved."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return np.ascontiguousarray(tiles[rank::world])


def reduce_film(film, dst: int = 0):
    """Sum-reduce a film tensor (torch, on the backend's device: NCCL for CUDA tensors, gloo for CPU) to `dst`."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film


def reduce_stats(values, dst: int = 0):
    """Sum a small vector of per-rank counters (ray counts, samples) to `dst`."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(values, dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
    return t
