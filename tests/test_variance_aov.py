"""AOV_OnlineMean / AOV_OnlineVariance: Welford's update as src/core/buffer/VarianceEstimator.inl:16-28 writes it,
fed with the per-iteration pixel value (FrameOutputDevice::mergeLocal, FrameOutputDevice.cpp:104-109).  The oracle is checked
against an independent float32 numpy restatement driven by per-iteration films; the device against the oracle, bit for bit."""
import numpy as np
import pytest

import pearray_b200 as prb
from oracle_binding import OracleScene
from scene_strings import MATERIAL_ZOO2

SCENE = MATERIAL_ZOO2.replace("(camera :name", "(filter :slot 'pixel' :type 'block' :radius 0)\n (output :name 'image' (channel :type 'color' :color 'xyz') (channel :type 'var' :color 'xyz') (channel :type 'online_mean' :color 'xyz'))\n (camera :name")
TILE = [(0, 0, 32, 32)]
ITERS = 6


def welford_restatement(values):
    """VarianceEstimator::addValue over the iterations 1..n, float32 like the reference"""
    f = np.float32
    mean = np.zeros_like(values[0], dtype=np.float32)
    var = np.zeros_like(mean)
    for i, v in enumerate(values, 1):
        delta = (v - mean).astype(np.float32)
        mean = (mean + (delta / f(i)).astype(np.float32)).astype(np.float32)
        delta2 = (v - mean).astype(np.float32)
        var = (((var * f(i - 1)).astype(np.float32) + (delta * delta2).astype(np.float32)).astype(np.float32) / f(i)).astype(np.float32)
    return mean, var


def test_scene_requests_the_estimator():
    scene = prb.Scene.from_string(SCENE)
    assert scene.settings.want_variance == 1 and scene.settings.filter_radius == 0
    assert prb.Scene.from_string(MATERIAL_ZOO2).settings.want_variance == 0


def test_oracle_variance_matches_independent_restatement():
    scene = prb.Scene.from_string(SCENE)
    ora = OracleScene(scene)
    # per-iteration pixel values: render iteration k alone into an empty film, carrying the RNG map along
    rng = scene.rng_map()
    values = []
    for k in range(ITERS):
        r = ora.render(TILE, k, 1, rng=rng, film=np.zeros((scene.height, scene.width, 3), np.float32), aov=False)
        rng = r["rng"]
        values.append(r["film"] * np.float32(k + 1))  # the mean after ONE fold into an empty film at iteration k+1 is x / (k+1)
    full = ora.render(TILE, 0, ITERS, rng=scene.rng_map(), aov=False, variance=True)
    mean, var = welford_restatement(values)
    # values recovered through x/(k+1)*(k+1) carry one rounding: compare with a tolerance far below the Monte-Carlo spread
    assert np.allclose(full["online_mean"], mean, rtol=2e-6, atol=1e-7)
    scale = max(1e-6, float(np.abs(var).max()))
    assert np.allclose(full["online_variance"], var, rtol=1e-4, atol=1e-6 * scale)
    assert full["online_variance"].max() > 0
    # the online mean is the plain mean of the samples (up to rounding): equals the film
    assert np.allclose(full["online_mean"], full["film"], rtol=1e-5, atol=1e-7)
    # resuming: 2 + 4 iterations == 6 iterations
    a = ora.render(TILE, 0, 2, rng=scene.rng_map(), aov=False, variance=True)
    b = ora.render(TILE, 2, 4, rng=a["rng"], film=a["film"], count=a["count"], aov=False, variance=(a["online_mean"], a["online_variance"]))
    assert np.array_equal(b["online_variance"].view(np.uint32), full["online_variance"].view(np.uint32))


@pytest.mark.gpu
def test_device_variance_bit_exact_and_written_to_the_output_file(tmp_path):
    import ctypes as C
    from test_image_output import read_exr
    scene = prb.Scene.from_string(SCENE)
    ctx = prb.Context(0)
    ctx.upload_scene(scene)
    ctx.upload_rng(scene.rng_map())
    ctx.render_tiles(TILE, 0, 2)
    ctx.render_tiles(TILE, 2, ITERS - 2)
    ref = OracleScene(scene).render(TILE, 0, ITERS, rng=scene.rng_map(), variance=True)
    mean, var = ctx.film_variance()
    assert np.array_equal(mean.view(np.uint32), ref["online_mean"].view(np.uint32))
    assert np.array_equal(var.view(np.uint32), ref["online_variance"].view(np.uint32))
    plain = prb.Context(0)
    plain.upload_scene(prb.Scene.from_string(MATERIAL_ZOO2))
    with pytest.raises(prb.PrbError):
        plain.film_variance()
    # through the host driver: results/image.exr carries variance.R/G/B and online_mean.R/G/B
    h = prb.host_lib()
    rc = h.prh_render_context_create(scene._h, 0, 0, 1)
    assert rc and h.prh_render_context_start(rc, 4, 4, ITERS) == 0
    assert h.prh_render_context_save_outputs(rc, str(tmp_path).encode()) == 1
    dev = h.prh_render_context_device(rc)
    v = np.empty((scene.height, scene.width, 3), np.float32)
    assert prb.device_lib().prb_film_download_variance(dev, None, v.ctypes.data_as(C.c_void_p)) == 0
    h.prh_render_context_destroy(rc)
    _, chans, pl = read_exr(str(tmp_path / "results" / "image.exr"))
    assert {"variance.R", "variance.G", "variance.B", "online_mean.R"} <= set(chans)
    assert np.array_equal(pl["variance.G"], v[..., 1]) and v.max() > 0
