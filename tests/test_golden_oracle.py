"""The oracle against its frozen outputs (tests/golden/*.npz, written by tools/make_golden.py) and end-to-end
properties of the restated path: the reference's furnace set-up and tile/GPU-partition invariance."""
import os

import numpy as np
import pytest

import pearray_b200 as prb
from conftest import ROOT, scene_path
from oracle_binding import OracleScene
from scene_strings import FURNACE, LPE_ZOO, MATERIAL_ZOO, MATERIAL_ZOO2, MATERIAL_ZOO3, MATERIAL_ZOO4, SKYSUN_ZOO

GOLDEN = ["c0_evaluation", "c1_sphere", "c2_cornellbox", "c3_cornellbox_glassy", "c4_boltsandgears", "c4b_complex_env", "c4c_complex", "material_zoo",
          "skysun_zoo", "material_zoo2", "material_zoo3", "material_zoo4", "lpe_zoo"]
STAT_NAMES = ["camera_ray_count", "light_ray_count", "primary_ray_count", "bounce_ray_count", "shadow_ray_count", "monochrome_ray_count",
              "pixel_sample_count", "entity_hit_count", "background_hit_count", "camera_depth_count", "light_depth_count"]


def load_scene(name):
    if name == "material_zoo":
        return prb.Scene.from_string(MATERIAL_ZOO)
    if name == "material_zoo3":
        return prb.Scene.from_string(MATERIAL_ZOO3)
    if name == "material_zoo4":
        return prb.Scene.from_string(MATERIAL_ZOO4)
    if name == "lpe_zoo":
        return prb.Scene.from_string(LPE_ZOO)
    if name == "material_zoo2":
        return prb.Scene.from_string(MATERIAL_ZOO2)
    if name == "skysun_zoo":
        return prb.Scene.from_string(SKYSUN_ZOO)
    return prb.Scene.from_file(scene_path(name + ".prc"))


def load_golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_hits_match_golden(name):
    g = load_golden(name)
    ora = OracleScene(load_scene(name))
    for tag in ("cam", "inc"):
        ent, prim, u, v, t = ora.trace_closest(g[tag + "_o"], g[tag + "_d"])
        assert np.array_equal(ent, g[tag + "_ent"]) and np.array_equal(prim, g[tag + "_prim"])
        for a, b in ((u, g[tag + "_u"]), (v, g[tag + "_v"]), (t, g[tag + "_t"])):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    occ = ora.trace_any(g["inc_o"], g["inc_d"], None, g["inc_tmax"])
    assert np.array_equal(occ, g["inc_occ"])
    # any-hit must agree with closest-hit restricted to the same interval
    assert np.array_equal(occ.astype(bool), (g["inc_ent"] != prb.INVALID_ID) & (g["inc_t"] <= g["inc_tmax"]))


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_film_matches_golden(name):
    g = load_golden(name)
    scene = load_scene(name)
    ora = OracleScene(scene)
    sx, sy, ex, ey = (int(x) for x in g["tile"])
    r = ora.render([(sx, sy, ex, ey)], 0, 4)
    assert np.array_equal(r["count"][sy:ey, sx:ex], g["count"])
    assert np.array_equal(r["rng"].reshape(scene.height, scene.width)[sy:ey, sx:ex], g["rng_after"])
    assert [r["stats"][k] for k in STAT_NAMES] == [int(x) for x in g["stats"]]
    assert np.array_equal(r["film"][sy:ey, sx:ex].view(np.uint32), g["film"].view(np.uint32))


def test_oracle_is_independent_of_tiling_and_threads():
    """a pixel only depends on its own RNG stream: any tiling / thread count gives the bit-identical film
    (the property the multi-GPU tile partition relies on)"""
    scene = load_scene("c3_cornellbox_glassy")
    ora = OracleScene(scene)
    region = (96, 96, 160, 160)
    a = ora.render([region], 0, 3, threads=1)
    sub = [(96, 96, 128, 160), (128, 96, 160, 120), (128, 120, 160, 160)]
    b = ora.render(sub, 0, 3, threads=4)
    assert np.array_equal(a["film"].view(np.uint32), b["film"].view(np.uint32))
    assert np.array_equal(a["rng"], b["rng"]) and a["stats"] == b["stats"]
    # resuming: 2 + 1 iterations == 3 iterations (running mean, FrameOutputDevice.cpp:202-221)
    c1 = ora.render([region], 0, 2, threads=2)
    c2 = ora.render([region], 2, 1, rng=c1["rng"], film=c1["film"], count=c1["count"], threads=2)
    assert np.array_equal(c2["film"].view(np.uint32), a["film"].view(np.uint32))


def test_furnace_hero():
    """reference furnace set-up (src/tests/python/whitefurnance.py: unit sphere, albedo 1, constant env radiance 1,
    direct depth 4).  Every path carries radiance 1 at every wavelength, so with hero wavelengths the XYZ film over the
    sphere converges to the XYZ of the flat unit spectrum (Y = 1) times the fraction of paths not cut at depth 4; a camera
    ray that misses everything is splatted with MIS = 1 for each of the 4 wavelengths (IntegratorUtils.h:16-53) -> 4x."""
    scene = prb.Scene.from_string(FURNACE % dict(hero="true"))
    ora = OracleScene(scene)
    r = ora.render([(0, 0, 48, 48)], 0, 64)
    f = r["filtered"]
    ys, xs = np.mgrid[0:48, 0:48]
    rad = np.hypot(xs - 23.5, ys - 23.5)
    inside = rad < 12  # well inside the sphere's silhouette (radius ~ 17 px)
    outside = rad > 22
    assert abs(f[inside][:, 1].mean() - 1.0) < 0.03
    assert abs(f[outside][:, 1].mean() - 4.0) < 0.12
    assert (r["count"][inside] == 64).all() and (r["count"][outside] == 0).all()


def test_furnace_non_hero_and_single_wavelength():
    """The other two modes of the reference furnace test (whitefurnance.py: `:spectral_hero false`, `:spectral_domain 520`).
    Both force monochrome camera rays (RenderTile.cpp:123-128): only the hero wavelength carries importance, so a camera ray
    that misses everything is splatted once instead of four times.  With a single-wavelength domain the sample is
    deterministic (wavelength 520 nm, pdf 1), so every background pixel holds the same value.  On the sphere the
    per-wavelength MIS weight of a monochrome NEE fragment divides by heroFactor = (1,0,0,0) (direct.cpp:318), its non-hero
    lanes are NaN and LocalFrameOutputDevice.cpp:128-144 drops the fragment: only BSDF-sampled background hits remain,
    the same fraction of the background value in both modes.  A single-wavelength render has a monotonic film
    (Environment.cpp:194-198): the unweighted hero sample in all three channels (LocalFrameOutputDevice.cpp:76-85), so the
    background is exactly the radiance 1.  (tests/test_whitefurnance.py is the literal port of the reference test.)"""
    ys, xs = np.mgrid[0:48, 0:48]
    rad = np.hypot(xs - 23.5, ys - 23.5)
    inside, outside = rad < 12, rad > 22
    nonhero = OracleScene(prb.Scene.from_string(FURNACE % dict(hero="false"))).render([(0, 0, 48, 48)], 0, 64)["filtered"]
    assert abs(nonhero[outside][:, 1].mean() - 1.0) < 0.05
    mono_src = (FURNACE % dict(hero="true")).replace(":spectral_hero true", ":spectral_hero true :spectral_domain 520")
    mono = OracleScene(prb.Scene.from_string(mono_src)).render([(0, 0, 48, 48)], 0, 64)["filtered"]
    bg = mono[outside]
    assert np.all(bg == 1.0)  # monotonic film: radiance 1, MIS 1, three equal channels
    r_mono = mono[inside][:, 1].mean() / bg[:, 1].mean()
    r_nonhero = nonhero[inside][:, 1].mean() / nonhero[outside][:, 1].mean()
    assert 0.05 < r_mono < 0.5 and abs(r_mono - r_nonhero) < 0.15 * r_mono


# ------------------------------------------------------------------ the reference's golden image
XYZ_TO_LINEAR_SRGB = np.array([[3.2404542, -1.5371385, -0.4985314], [-0.9692660, 1.8760108, 0.0415560], [0.0556434, -0.2040259, 1.0572252]], np.float32)
CBOX_LUMINANCE_TOL = 0.2   # relative RMSE of the 8x8-block luminance outside the luminaire
CBOX_CHANNEL_TOL = 0.2     # relative difference of the per-channel image means


def cbox_reference_error(xyz_film):
    """Compare a render of scenes/c0_evaluation.prc with examples/evaluation/cbox.exr (tests/golden/cbox_reference_blocks.npz,
    tools/make_golden.py:import_reference_image).  The image comes from another renderer (Mitsuba 2: CIE 1931 observer,
    360-830 nm, different path-depth convention), so the bar is a sanity bound on block means, not bit parity."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "cbox_reference_blocks.npz"))
    rgb = xyz_film.astype(np.float32) @ XYZ_TO_LINEAR_SRGB.T
    w = g["pixel_mask"].astype(np.float32)[..., None]
    blocks = (rgb * w).reshape(32, 8, 32, 8, 3).sum(axis=(1, 3)) / np.maximum(w.reshape(32, 8, 32, 8, 1).sum(axis=(1, 3)), 1)
    lum = np.array([0.2126, 0.7152, 0.0722], np.float32)
    a, b = blocks @ lum, g["blocks"] @ lum
    lum_err = float(np.sqrt(np.mean((a - b) ** 2)) / np.mean(b))
    ratio = blocks.sum(axis=(0, 1)) / g["blocks"].sum(axis=(0, 1))
    return lum_err, ratio


def test_oracle_vs_reference_golden_image():
    scene = load_scene("c0_evaluation")
    ora = OracleScene(scene)
    r = ora.render(scene.tiles(8, 8), 0, 32, rng=scene.rng_map(), threads=os.cpu_count() or 1, aov=False)
    lum_err, ratio = cbox_reference_error(r["filtered"])
    assert lum_err < CBOX_LUMINANCE_TOL, lum_err
    assert np.all(np.abs(ratio - 1) < CBOX_CHANNEL_TOL), ratio
