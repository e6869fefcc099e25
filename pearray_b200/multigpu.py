"""Multi-GPU partitioning of the render: scene replicated per GPU, one process per GPU.

Two partitions of the work, both ending in ONE reduction of the per-rank films (torch.distributed; NCCL over NVLink
on the GPU box, gloo in the CPU tests):

  tiles    render tiles of RenderTileMap (reference src/core/renderer/RenderTileMap.cpp:26-122) are owned interleaved,
           tile_id % world == rank.  A pixel's RNG stream and film cell belong to exactly one rank, so the reduced
           film is BIT-IDENTICAL to the single-GPU film for any world size (non-owned cells are zero -> plain sum).
  samples  every rank renders the whole film for its own block of iterations from a decorrelated RNG map (seed +
           7919 * rank); the reduced film is the average of the rank films.  Statistically equivalent, not bit-identical
           (a pixel's PCG stream is consumed with a data-dependent number of draws, it cannot be split by jump-ahead).

The pixel filter is linear with integer offsets, so it is applied after the reduce (unfiltered films are summed).
"""
import numpy as np

SEED_STRIDE = 7919


def partition_tiles(tiles, rank, world):
    """interleaved tile ownership"""
    return [t for i, t in enumerate(tiles) if i % world == rank]


def rank_seed(seed, rank):
    return int(seed) + SEED_STRIDE * int(rank)


def pack_film(xyz_unfiltered, count):
    """(H,W,3) float32 mean + (H,W) uint32 sample count -> (H,W,4) float32 buffer handed to the reduce (same layout as
    prb_film_export_device)"""
    out = np.empty(xyz_unfiltered.shape[:2] + (4,), np.float32)
    out[..., :3] = xyz_unfiltered
    out[..., 3] = count
    return out


def reduce_film(film4, mode, world, dst=0):
    """film4: torch tensor (..., 4) [x, y, z, count] of this rank.  In-place reduce to `dst`; returns the tensor.
    mode 'tiles': plain sum.  mode 'samples': xyz averaged over ranks, counts summed."""
    import torch.distributed as dist
    if world == 1:
        return film4
    if mode == "samples":
        film4[..., :3] /= world
    elif mode != "tiles":
        raise ValueError("mode must be 'tiles' or 'samples'")
    dist.reduce(film4, dst=dst, op=dist.ReduceOp.SUM)
    return film4
