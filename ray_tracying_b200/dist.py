"""Screen-tile sharding across GPUs: one process per GPU, scene and BVH replicated, rank r renders
the screen tiles t with t % world == r (rt_render_params.rank/world), and the frame is assembled
once at frame end -- there is no inter-GPU traffic inside the render loop.

The reference renders the whole frame in one serial loop (raytracer.cpp:433-476); pixels are
independent, so any partition of them gives the same image. With the counter-based RNG the image
does not depend on the partition either (tests/test_gpu_parity.py checks that bit for bit).

torch / torch.distributed are plumbing here: index tensors, one all_gather of each rank's packed
tile pixels (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def tile_grid(width: int, height: int, tile: Tuple[int, int]) -> Tuple[int, int]:
    return (width + tile[0] - 1) // tile[0], (height + tile[1] - 1) // tile[1]


def tile_owner(width: int, height: int, tile: Tuple[int, int], world: int) -> np.ndarray:
    """(H, W) int32: rank that renders each pixel (same rule as finalize_kernel in render.cu)."""
    tiles_x, _ = tile_grid(width, height, tile)
    ys, xs = np.mgrid[0:height, 0:width]
    t = (ys // tile[1]) * tiles_x + (xs // tile[0])
    return (t % world).astype(np.int32)


def rank_pixel_indices(width: int, height: int, tile: Tuple[int, int], world: int, rank: int) -> np.ndarray:
    """Flat pixel indices (row-major) owned by `rank`, in increasing order."""
    return np.flatnonzero(tile_owner(width, height, tile, world).reshape(-1) == rank).astype(np.int64)


def max_rank_pixels(width: int, height: int, tile: Tuple[int, int], world: int) -> int:
    owner = tile_owner(width, height, tile, world).reshape(-1)
    return int(np.bincount(owner, minlength=world).max())


def gather_frame(local_frame, width: int, height: int, tile: Tuple[int, int], rank: int, world: int, group=None):
    """Assembles the full frame on every rank.

    local_frame: torch tensor (H, W, C) on this rank's device in which only this rank's tiles are
    valid. Each rank packs its own pixels, one all_gather moves world x max_rank_pixels x C
    elements, and every rank scatters the pieces into a full (H, W, C) tensor.
    """
    import torch
    import torch.distributed as dist

    if world == 1:
        return local_frame
    c = local_frame.shape[-1]
    dev = local_frame.device
    flat = local_frame.reshape(-1, c)
    cap = max_rank_pixels(width, height, tile, world)
    mine = torch.from_numpy(rank_pixel_indices(width, height, tile, world, rank)).to(dev)
    packed = torch.zeros((cap, c), dtype=local_frame.dtype, device=dev)
    packed[: mine.numel()] = flat.index_select(0, mine)
    gathered = torch.empty((world, cap, c), dtype=local_frame.dtype, device=dev)
    dist.all_gather_into_tensor(gathered.view(-1, c), packed, group=group)
    out = torch.empty_like(flat)
    for r in range(world):
        idx = mine if r == rank else torch.from_numpy(rank_pixel_indices(width, height, tile, world, r)).to(dev)
        out.index_copy_(0, idx, gathered[r, : idx.numel()])
    return out.reshape(height, width, c)
