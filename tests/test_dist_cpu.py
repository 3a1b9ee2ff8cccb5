"""CPU tests of the multi-GPU host logic with torch.distributed (gloo, world_size 2 and 3):
tile ownership, packing, the frame-end all_gather and reassembly. The 'render' is replaced by a
synthetic per-pixel function so no GPU is needed; on GPUs the same gather_frame() runs over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ray_tracying_b200 import dist as rdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _pixel_function(width, height):
    ys, xs = np.mgrid[0:height, 0:width]
    return np.stack([(xs * 7 + ys * 3) % 251, (xs + ys * 5) % 241, (xs * ys) % 239], axis=-1).astype(np.uint8)


def _worker(rank, world, port, width, height, tile, out_dir, block=0):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        truth = _pixel_function(width, height)
        owner = rdist.tile_owner(width, height, tile, world, block)
        local = np.zeros_like(truth)
        local[owner == rank] = truth[owner == rank]  # what rt_render_device leaves in a rank's frame
        frame = rdist.gather_frame(torch.from_numpy(local), width, height, tile, rank, world, block=block)
        np.save(os.path.join(out_dir, f"frame_{rank}.npy"), frame.numpy())
        # ids travel the same way (int32, one channel)
        ids_truth = (np.arange(width * height, dtype=np.int32).reshape(height, width) % 1000) - 1
        ids_local = np.full_like(ids_truth, -1)
        ids_local[owner == rank] = ids_truth[owner == rank]
        ids = rdist.gather_frame(torch.from_numpy(ids_local)[..., None], width, height, tile, rank, world, block=block)[..., 0]
        assert np.array_equal(ids.numpy(), ids_truth)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,width,height,tile,block", [(2, 100, 60, (32, 32), 0), (3, 97, 45, (16, 8), 0), (2, 64, 36, (8, 4), 0),
                                                         (2, 100, 60, (8, 4), 3)])
def test_gather_frame_gloo(tmp_path, world, width, height, tile, block):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, width, height, tile, str(tmp_path), block), nprocs=world, join=True)
    truth = _pixel_function(width, height)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"frame_{r}.npy"), truth)


def test_tile_owner_interleaves_tiles():
    owner = rdist.tile_owner(100, 60, (32, 32), 4)
    assert owner.shape == (60, 100)
    assert owner[0, 0] == 0 and owner[0, 32] == 1 and owner[0, 64] == 2 and owner[0, 96] == 3
    assert owner[32, 0] == 0  # 4 tiles per row -> second tile row starts at tile 4
    counts = np.bincount(owner.reshape(-1), minlength=4)
    assert counts.sum() == 6000 and rdist.max_rank_pixels(100, 60, (32, 32), 4) == counts.max()
    idx = rdist.rank_pixel_indices(100, 60, (32, 32), 4, 2)
    assert np.all(owner.reshape(-1)[idx] == 2) and len(idx) == counts[2]


def test_tile_owner_in_blocks():
    """block = B: B x B groups of tiles go to one rank (render.cu tile_owner): 8x4 tiles, blocks of 2 -> 16x8 pixel groups."""
    owner = rdist.tile_owner(64, 32, (8, 4), 3, block=2)
    assert (owner[:8, :16] == 0).all() and (owner[:8, 16:32] == 1).all() and (owner[:8, 32:48] == 2).all() and (owner[:8, 48:64] == 0).all()
    assert (owner[8:16, :16] == 1).all()  # 4 blocks per row of blocks: the second row starts at block 4 -> rank 1
