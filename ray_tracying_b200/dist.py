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


def tile_owner(width: int, height: int, tile: Tuple[int, int], world: int, block: int = 0) -> np.ndarray:
    """(H, W) int32: rank that renders each pixel (same rule as tile_owner() in render.cu): tiles are
    dealt round-robin in row-major order, singly or -- block = B > 1 -- in B x B groups."""
    tiles_x, _ = tile_grid(width, height, tile)
    ys, xs = np.mgrid[0:height, 0:width]
    tx, ty = xs // tile[0], ys // tile[1]
    if block <= 1:
        t = ty * tiles_x + tx
    else:
        t = (ty // block) * ((tiles_x + block - 1) // block) + tx // block
    return (t % world).astype(np.int32)


def rank_pixel_indices(width: int, height: int, tile: Tuple[int, int], world: int, rank: int, block: int = 0) -> np.ndarray:
    """Flat pixel indices (row-major) owned by `rank`, in increasing order."""
    return np.flatnonzero(tile_owner(width, height, tile, world, block).reshape(-1) == rank).astype(np.int64)


def max_rank_pixels(width: int, height: int, tile: Tuple[int, int], world: int, block: int = 0) -> int:
    owner = tile_owner(width, height, tile, world, block).reshape(-1)
    return int(np.bincount(owner, minlength=world).max())


_PLAN_CACHE: dict = {}


def _gather_plan(width: int, height: int, tile: Tuple[int, int], world: int, rank: int, device, block: int = 0):
    """Index tensors of the frame-end exchange, built once per (frame geometry, world, device):
    `mine` = this rank's pixels (for packing), `scatter` = for every slot of the gathered
    (world x cap) buffer the flat pixel it belongs to (padding slots point at a scratch pixel)."""
    import torch

    key = (width, height, tuple(tile), world, rank, str(device), block)
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        owner = tile_owner(width, height, tile, world, block).reshape(-1)
        cap = int(np.bincount(owner, minlength=world).max())
        scatter = np.full((world, cap), width * height, dtype=np.int64)  # scratch slot = one past the frame
        mine = None
        for r in range(world):
            idx = np.flatnonzero(owner == r).astype(np.int64)
            scatter[r, : idx.size] = idx
            if r == rank:
                mine = idx
        plan = (torch.from_numpy(mine).to(device), torch.from_numpy(scatter.reshape(-1)).to(device), cap)
        _PLAN_CACHE[key] = plan
    return plan


_BUFFERS: dict = {}


def _buffer(name: str, shape, dtype, device):
    """Exchange buffers are kept per (shape, dtype, device): a frame allocates nothing."""
    import torch

    key = (name, tuple(shape), dtype, str(device))
    buf = _BUFFERS.get(key)
    if buf is None:
        buf = torch.zeros(shape, dtype=dtype, device=device)
        _BUFFERS[key] = buf
    return buf


def gather_frame(local_frame, width: int, height: int, tile: Tuple[int, int], rank: int, world: int, group=None, block: int = 0):
    """Assembles the full frame on every rank.

    local_frame: torch tensor (H, W, C) on this rank's device in which only this rank's tiles are
    valid. Each rank packs its own pixels (one gather kernel), ONE all_gather moves
    world x max_rank_pixels x C elements over NCCL / NVLink, and one scatter kernel writes the
    pieces into a full (H, W, C) tensor. The index tensors and the three exchange buffers are cached
    per frame geometry, so the returned frame is overwritten by the next call with the same geometry.
    """
    import torch
    import torch.distributed as dist

    if world == 1:
        return local_frame
    c = local_frame.shape[-1]
    dev = local_frame.device
    flat = local_frame.reshape(-1, c)
    mine, scatter, cap = _gather_plan(width, height, tile, world, rank, dev, block)
    packed = _buffer("packed", (cap, c), local_frame.dtype, dev)  # the padding rows past mine.numel() stay zero
    torch.index_select(flat, 0, mine, out=packed[: mine.numel()])
    gathered = _buffer("gathered", (world * cap, c), local_frame.dtype, dev)
    dist.all_gather_into_tensor(gathered, packed, group=group)
    out = _buffer("frame", (width * height + 1, c), local_frame.dtype, dev)  # + scratch pixel
    out.index_copy_(0, scatter, gathered)
    return out[: width * height].reshape(height, width, c)
