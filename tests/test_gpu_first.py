"""First GPU parity checks: primary-ray hit ids bit exact vs the oracle, film within tolerance."""
import numpy as np
import pytest

import pearray_b200 as prb
from conftest import scene_path
from oracle_binding import OracleScene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cbox():
    scene = prb.Scene.from_file(scene_path("c2_cornellbox.prc"))
    ctx = prb.Context(0)
    ctx.upload_scene(scene)
    ctx.upload_rng(scene.rng_map())
    return scene, ctx, OracleScene(scene)


def test_camera_rays_bit_exact(cbox):
    scene, ctx, ora = cbox
    tiles = [(0, 0, 500, 500)]
    org, dr, wvl, pix = ctx.generate_camera_rays(tiles, 3)
    oorg, odr, owvl, opix = ora.generate_camera_rays(tiles, 3)
    assert np.array_equal(pix, opix)
    assert np.array_equal(org.view(np.uint32), oorg.view(np.uint32))
    assert np.array_equal(dr.view(np.uint32), odr.view(np.uint32))
    assert np.array_equal(wvl.view(np.uint32), owvl.view(np.uint32))


def test_primary_hits_bit_exact(cbox):
    scene, ctx, ora = cbox
    org, dr, wvl, pix = ctx.generate_camera_rays([(0, 0, 500, 500)], 0)
    got = ctx.trace_closest(org, dr)
    ref = ora.trace_closest(org, dr)
    assert np.array_equal(got[0], ref[0]), "entity ids"
    assert np.array_equal(got[1], ref[1]), "primitive ids"
    for a, b in zip(got[2:], ref[2:]):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_film_matches_oracle(cbox):
    scene, ctx, ora = cbox
    tiles = [(128, 128, 256, 256)]
    ctx.upload_rng(scene.rng_map())
    ctx.film_clear()
    ctx.reset_stats()
    ctx.render_tiles(tiles, 0, 8)
    xyz, cnt = ctx.film()
    ref = ora.render(tiles, 0, 8)
    a = xyz[128:256, 128:256].astype(np.float64)
    b = ref["filtered"][128:256, 128:256].astype(np.float64)
    rel = np.sqrt(np.mean((a - b) ** 2)) / np.mean(b)
    frac_bad = np.mean(np.abs(a - b) > 1e-4 * (np.abs(b) + 1e-3))
    print("relRMSE", rel, "frac pixels differing", frac_bad)
    assert np.array_equal(cnt, ref["count"])
    assert rel < 1e-2
    st = ctx.stats()
    for k, v in ref["stats"].items():
        g = getattr(st, k)
        assert abs(int(g) - v) <= max(4, 1e-3 * v), (k, int(g), v)
    assert np.array_equal(ctx.download_rng()[cnt.reshape(-1) > 0], ref["rng"][cnt.reshape(-1) > 0]) or frac_bad > 0
