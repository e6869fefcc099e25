"""world_size-2 CPU test (gloo) of the multi-GPU host logic: interleaved tile ownership + ONE reduce of the
per-rank films gives the bit-identical film of a single-rank render; the sample-range mode averages rank films.
The per-rank "device" here is the oracle (no GPU in this container); the partition / reduce code is the product's
(pearray_b200/multigpu.py), the same functions bench.py uses with NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pearray_b200 as prb
from conftest import ROOT, scene_path

REGION = (100, 100, 164, 164)
ITER = 2


def _worker(rank, world, port, mode, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_binding import OracleScene
    from pearray_b200 import multigpu
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = prb.Scene.from_file(scene_path("c3_cornellbox_glassy.prc"))
    sx, sy, ex, ey = REGION
    tiles = [(x, y, x + 16, y + 16) for y in range(sy, ey, 16) for x in range(sx, ex, 16)]
    if mode == "tiles":
        mine = multigpu.partition_tiles(tiles, rank, world)
    else:
        mine = tiles
        scene.settings.seed = multigpu.rank_seed(scene.settings.seed, rank)
    ora = OracleScene(scene)
    r = ora.render(mine, 0, ITER, threads=2)
    film4 = torch.from_numpy(multigpu.pack_film(r["film"], r["count"]))
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), film4.numpy().copy())
    multigpu.reduce_film(film4, mode, world, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), film4.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _run(mode, tmp_path, port):
    mp.spawn(_worker, args=(2, port, mode, str(tmp_path)), nprocs=2, join=True)
    return np.load(tmp_path / "reduced.npy"), [np.load(tmp_path / ("rank%d.npy" % r)) for r in range(2)]


def test_tile_partition_is_bit_identical(tmp_path):
    from oracle_binding import OracleScene
    from pearray_b200 import multigpu
    reduced, ranks = _run("tiles", tmp_path, 29611)
    scene = prb.Scene.from_file(scene_path("c3_cornellbox_glassy.prc"))
    single = OracleScene(scene).render([REGION], 0, ITER, threads=4)
    ref = multigpu.pack_film(single["film"], single["count"])
    assert np.array_equal(reduced.view(np.uint32), ref.view(np.uint32))
    # ownership is disjoint: no film cell written by both ranks
    assert not ((ranks[0][..., 3] > 0) & (ranks[1][..., 3] > 0)).any()


def test_sample_partition_averages_rank_films(tmp_path):
    reduced, ranks = _run("samples", tmp_path, 29612)
    assert np.allclose(reduced[..., :3], 0.5 * (ranks[0][..., :3] + ranks[1][..., :3]), rtol=1e-6, atol=1e-7)
    assert np.array_equal(reduced[..., 3], ranks[0][..., 3] + ranks[1][..., 3])
    assert not np.array_equal(ranks[0], ranks[1])  # decorrelated RNG maps


def test_partition_tiles_interleaves():
    from pearray_b200 import multigpu
    tiles = list(range(10))
    parts = [multigpu.partition_tiles(tiles, r, 4) for r in range(4)]
    assert sorted(sum(parts, [])) == tiles and parts[1] == [1, 5, 9]
    with pytest.raises(ValueError):
        multigpu.reduce_film(torch.zeros(2, 4), "bogus", 2)
