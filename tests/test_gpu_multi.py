"""Multi-GPU tests (`pytest -m gpu` on a box with >= 2 GPUs; skipped otherwise): one process per GPU, the film combine of
the C ABI over NCCL (prb_comm_init + prb_film_reduce_comm), checked against a single-GPU render.  The unique id travels
through a file, as a C++ client without torch would ship it."""
import os
import sys

import numpy as np
import pytest

import pearray_b200 as prb
from conftest import ROOT, scene_path

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        return prb.device_lib().prb_device_count()
    except Exception:
        return 0


def _worker(rank, world, out_dir, partition, scene_name, spp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import time
    from pearray_b200 import multigpu
    idf = os.path.join(out_dir, "nccl_id.bin")
    if rank == 0:
        uid = prb.Context.comm_unique_id()
        with open(idf + ".tmp", "wb") as f:
            f.write(uid.tobytes())
        os.replace(idf + ".tmp", idf)
    else:
        for _ in range(600):
            if os.path.exists(idf):
                break
            time.sleep(0.05)
        uid = np.frombuffer(open(idf, "rb").read(), np.uint8)
    scene = prb.Scene.from_file(scene_path(scene_name))
    tiles = scene.tiles(8, 8)
    if partition == "tiles":
        mine, first = multigpu.partition_tiles(tiles, rank, world), 0
    else:
        mine, first = tiles, rank * spp
        scene.settings.seed = multigpu.rank_seed(scene.settings.seed, rank)
    ctx = prb.Context(rank)
    ctx.upload_scene(scene)
    ctx.upload_rng(scene.rng_map())
    ctx.comm_init(uid, rank, world)
    ctx.render_tiles(mine, first, spp)
    own, own_cnt = ctx.film()
    np.save(os.path.join(out_dir, "own%d.npy" % rank), own)
    ctx.film_reduce_comm(partition, spp * world if partition == "samples" else spp, 0)
    if rank == 0:
        xyz, cnt = ctx.film()
        np.save(os.path.join(out_dir, "reduced.npy"), xyz)
        np.save(os.path.join(out_dir, "reduced_cnt.npy"), cnt)
        np.save(os.path.join(out_dir, "reduced_aov.npy"), ctx.film_aov())
    ctx.comm_destroy()


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_tile_partition_bit_identical(tmp_path, world):
    """strong-scaling mode: interleaved tiles over `world` GPUs, ONE ncclReduce -> the film on rank 0 is bit-identical to the
    single-GPU render (xyz through the pixel filter, sample counts, AOV sums)"""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    spp = 3
    mp.spawn(_worker, args=(world, str(tmp_path), "tiles", "c3_cornellbox_glassy.prc", spp), nprocs=world, join=True)
    scene = prb.Scene.from_file(scene_path("c3_cornellbox_glassy.prc"))
    ctx = prb.Context(0)
    ctx.upload_scene(scene)
    ctx.upload_rng(scene.rng_map())
    ctx.render_tiles(scene.tiles(8, 8), 0, spp)
    xyz, cnt = ctx.film()
    assert np.array_equal(np.load(tmp_path / "reduced.npy").view(np.uint32), xyz.view(np.uint32))
    assert np.array_equal(np.load(tmp_path / "reduced_cnt.npy"), cnt)
    assert np.array_equal(np.load(tmp_path / "reduced_aov.npy").view(np.uint32), ctx.film_aov().view(np.uint32))


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_nccl_sample_ranges(tmp_path):
    """weak-scaling mode: rank r renders iterations [r spp, (r+1) spp) of one sequence; the reduced film is the mean over all"""
    import torch.multiprocessing as mp
    spp, world = 4, 2
    mp.spawn(_worker, args=(world, str(tmp_path), "samples", "c0_evaluation.prc", spp), nprocs=world, join=True)
    own = [np.load(tmp_path / ("own%d.npy" % r)).astype(np.float64) for r in range(world)]
    expect = sum(own[r] * ((r + 1) * spp / (world * spp)) for r in range(world))
    assert np.allclose(np.load(tmp_path / "reduced.npy"), expect, rtol=1e-6, atol=1e-9)
