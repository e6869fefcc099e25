"""An INDEPENDENT restatement of the MIS weights of the `direct` integrator, written in float64 numpy from the text of the
reference (src/plugins/main/integrators/direct.cpp:233-456, src/vcm/vcm/MIS.h:7-30, IntegratorUtils.h:16-53) and of the
film commit (src/loader/output/LocalFrameOutputDevice.cpp:88-164, src/core/spectral/CIE.h:30-62) -- NOT from oracle.cpp.

The oracle logs every fragment it pushes together with the path state its weight was computed from (orc_log_fragments,
test instrumentation); here the weight is recomputed from that state and the film pixel is rebuilt from the fragments.
This breaks the "same text twice" symmetry between oracle.cpp and the device code: a transcription error of direct.cpp's
MIS / hero-wavelength handling made in the oracle (and copied into the device code) shows up here.

The furnace test below pins the remaining links of the chain -- light sampling pdf, material pdf, environment evaluation
and the CIE normalisation -- against a closed-form expectation evaluated by quadrature."""
import os

import numpy as np
import pytest

import pearray_b200 as prb
from conftest import scene_path
from oracle_binding import OracleScene
from scene_strings import MATERIAL_ZOO, SKYSUN_ZOO, WHITEFURNACE_FULL

FK_BACKGROUND, FK_DIRECT_HIT, FK_NEE, FK_INF_LIGHT, FK_ZERO = range(5)
FF_RAY_MONO, FF_BSDF_MONO, FF_LIGHT_DELTA, FF_LAST_DELTA, FF_LAST_EMISSIVE, FF_FROM_BEHIND, FF_VISIBLE, FF_LIGHT_INFINITE = (1 << i for i in range(8))
HERO_ONLY = np.array([1.0, 0.0, 0.0, 0.0])
ONES = np.ones(4)


def mis_term(power, a):
    """VCM::mis_term<MISMode>, vcm/MIS.h:7-30: balance -> a, power -> a^2"""
    return a * a if power else a


def restated_mis(rec, i, settings):
    """the weight direct.cpp hands to pushSpectralFragment for record i, from the reference text"""
    kind, flags = int(rec["kind"][i]), int(rec["flags"][i])
    power, do_nee, do_direct = bool(settings.mis_power), bool(settings.do_nee), bool(settings.do_direct)
    path, prev, wvl = (rec[k][i].astype(np.float64) for k in ("pathPDF", "prevPathPDF", "wvlPDF"))
    ray_hero = HERO_ONLY if flags & FF_RAY_MONO else ONES
    with np.errstate(divide="ignore", invalid="ignore"):
        if kind == FK_BACKGROUND:  # IntegratorUtils.h:42: pushSpectralFragment(Ones, Ones, radiance, ...)
            return ONES
        if kind == FK_ZERO:  # direct.cpp:459-464
            return ray_hero / (ray_hero.sum() * wvl)
        if kind == FK_DIRECT_HIT:  # direct.cpp:355-412
            if not do_nee or flags & FF_FROM_BEHIND or flags & FF_LAST_DELTA:
                return ray_hero / (ray_hero.sum() * wvl)
            pos_pdf_s = float(rec["lightPdfS"][i])
            denom = mis_term(power, prev * ONES * pos_pdf_s).sum() + mis_term(power, path).sum()
            return ray_hero * mis_term(power, path[0]) / (denom * mis_term(power, wvl))
        if kind == FK_INF_LIGHT:  # direct.cpp:415-456
            if not do_nee or flags & FF_LAST_DELTA:
                return ray_hero / (ray_hero.sum() * wvl)
            denom_mis = sum(mis_term(power, prev * ONES * float(p)).sum() for p in rec["infPdfS"][i][:int(rec["extra"][i])])
            denom = mis_term(power, path).sum() + denom_mis
            return ray_hero * mis_term(power, path[0]) / (denom * mis_term(power, wvl))
        assert kind == FK_NEE  # direct.cpp:272-326
        hero = HERO_ONLY if flags & FF_BSDF_MONO else ONES
        bsdf_wvl_pdf = rec["bsdfPDF"][i].astype(np.float64) * hero
        light_pdf = 1.0 if flags & FF_LIGHT_DELTA else float(rec["lightPdfS"][i])
        light_pdf2 = light_pdf * ONES * ray_hero
        if do_direct and not flags & FF_LAST_EMISSIVE:
            bsdf_pdf = bsdf_wvl_pdf * float(rec["roulette"][i])
            denom = mis_term(power, path * light_pdf2).sum() + mis_term(power, path * bsdf_pdf).sum()
            if flags & FF_LIGHT_DELTA:
                return hero / hero.sum()
            return mis_term(power, path[0] * light_pdf2[0]) / (hero * denom * mis_term(power, wvl))
        return hero / (hero.sum() * wvl)


def same(a, b, rtol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    fin = np.isfinite(a) & np.isfinite(b)
    if not np.array_equal(np.isnan(a), np.isnan(b)) or not np.array_equal(np.isinf(a), np.isinf(b)):
        return False
    if not np.array_equal(np.sign(a[~fin & ~np.isnan(a)]), np.sign(b[~fin & ~np.isnan(b)])):
        return False
    return np.allclose(a[fin], b[fin], rtol=rtol, atol=1e-30)


def cie_xyz(wavelength):
    """CIE::eval(wavelength): table lookup / Y_NORM * RANGE (CIE.h:44-62) through the host library's table probe"""
    return np.array([prb.host_lib().prh_cie_eval(c, float(wavelength)) for c in range(3)], np.float64)


def restated_film(rec, monotonic, n_iterations):
    """LocalFrameOutputDevice::commitSpectrals2 + FrameOutputDevice::onEndOfIteration for the logged pixels: the running
    mean over iterations of sum_fragments CIE(heroFactor * MIS * Importance * Radiance) (radius-0 filter, BlendWeight 1)"""
    film = {}
    for i in range(len(rec["kind"])):
        mono = monotonic or (int(rec["flags"][i]) & FF_RAY_MONO)
        hero = HERO_ONLY if mono else ONES
        with np.errstate(invalid="ignore", over="ignore"):
            contrib = hero * (rec["mis"][i].astype(np.float64) * rec["importance"][i] * rec["radiance"][i])
        invalid = np.isnan(contrib).any() or np.isinf(contrib).any() or (contrib < -1e-5).any()
        assert (rec["accepted"][i] > 0) == (not invalid)
        if invalid:
            continue
        if monotonic:
            xyz = np.full(3, contrib[0])
        else:
            xyz = sum(contrib[k] * cie_xyz(rec["wvl"][i][k]) for k in range(4))
        pix = int(rec["pixel"][i])
        film[pix] = film.get(pix, np.zeros(3)) + xyz
    return {p: v / n_iterations for p, v in film.items()}


CASES = [("c2_cornellbox", 2), ("c3_cornellbox_glassy", 4), ("c4c_complex", 2), ("c1_sphere", 2), ("zoo", 4), ("skysun", 4), ("furnace_nonhero", 2), ("furnace_mono", 2)]


def load(name):
    if name == "zoo":
        return prb.Scene.from_string(MATERIAL_ZOO)
    if name == "skysun":
        return prb.Scene.from_string(SKYSUN_ZOO)
    if name == "furnace_nonhero":
        return prb.Scene.from_string(WHITEFURNACE_FULL.format(hero="false", size=32))
    if name == "furnace_mono":
        return prb.Scene.from_string(WHITEFURNACE_FULL.format(hero="true", size=32).replace(":spectral_hero true", ":spectral_hero true :spectral_domain 520"))
    return prb.Scene.from_file(scene_path(name + ".prc"))


@pytest.mark.parametrize("name,iters", CASES)
def test_mis_weights_and_film_against_independent_restatement(name, iters):
    scene = load(name)
    ora = OracleScene(scene)
    w, h = scene.width, scene.height
    rs = np.random.RandomState(7)
    n_pix = 160
    pixels = np.unique((rs.randint(0, h, n_pix) * w + rs.randint(0, w, n_pix)).astype(np.uint32))
    rng = scene.rng_map()
    rec = ora.log_fragments(pixels, 0, iters, rng=rng)
    n = len(rec["kind"])
    assert n >= len(pixels)
    kinds = set()
    for i in range(n):
        kinds.add(int(rec["kind"][i]))
        expect = restated_mis(rec, i, scene.settings)
        assert same(rec["mis"][i], expect, 2e-5), (name, i, int(rec["kind"][i]), int(rec["flags"][i]), rec["mis"][i], expect)
    # the film of the same pixels rebuilt from the fragments (block filter scenes only: the reference splats filtered fragments)
    r = ora.render([(int(p % w), int(p // w), int(p % w) + 1, int(p // w) + 1) for p in pixels], 0, iters, rng=rng, threads=1, aov=False)
    film = restated_film(rec, bool(scene.settings.film_monotonic), iters)
    got = r["film"].reshape(-1, 3)
    for p in pixels:
        e = film.get(int(p), np.zeros(3))
        assert np.allclose(got[p], e, rtol=2e-4, atol=1e-6 * max(1.0, float(np.abs(e).max()))), (name, int(p), got[p], e)



def test_every_fragment_kind_is_covered():
    kinds = set()
    for name, iters in CASES:
        scene = load(name)
        step, iters = (1, 8) if name == "skysun" else (37, 1)  # the delta sun is selected rarely
        rec = OracleScene(scene).log_fragments(np.arange(0, scene.width * scene.height, step, dtype=np.uint32)[:2304 if step == 1 else 200], 0, iters)
        kinds |= set(int(k) for k in rec["kind"])
        flags = np.bitwise_or.reduce(rec["flags"].astype(np.uint32)) if len(rec["flags"]) else 0
        if name == "c3_cornellbox_glassy":
            assert flags & FF_LAST_DELTA and flags & FF_RAY_MONO  # delta glass + hero collapse (dispersion)
        if name == "skysun":
            assert flags & FF_LIGHT_DELTA and flags & FF_LIGHT_INFINITE
    assert kinds >= {FK_BACKGROUND, FK_DIRECT_HIT, FK_NEE, FK_INF_LIGHT}


# ------------------------------------------------------------------ closed-form furnace
def furnace_expectation(normal, n_theta=400, n_phi=800):
    """Unit-albedo Lambert point with normal N under the constant environment of radiance 1 (no occluder), direct.cpp with NEE +
    BSDF sampling, balance heuristic, four hero wavelengths with the `random` mapper (wavelength pdf 1, CIE::eval carries
    RANGE / Y_NORM so that E[y(lambda)] = 1):
      NEE   samples L with the environment's pdf p_l = cos_hemi_pdf(L.z) on the UPPER hemisphere only (environment.cpp:92-93),
            weight per wavelength p_l / (4 p_l + 4 p_b) (direct.cpp:314-318), four wavelengths
            -> Y_nee  = int_{L.z > 0, N.L > 0} (N.L / pi) * p_l / (p_l + p_b) dL
      BSDF  samples L with p_b = N.L / pi, escapes, handleInfLights weight p_b / (4 p_b + 4 p_l'), p_l' = cos_hemi_pdf(|L.z|)
            (environment.cpp:75: the absolute value makes the lower hemisphere count in the pdf although it is never sampled)
            -> Y_bsdf = int_{N.L > 0} p_b * p_b / (p_b + p_l') dL
    evaluated by midpoint quadrature over the sphere in float64."""
    th = (np.arange(n_theta) + 0.5) * np.pi / n_theta
    ph = (np.arange(n_phi) + 0.5) * 2 * np.pi / n_phi
    T, P = np.meshgrid(th, ph, indexing="ij")
    L = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], -1)
    dw = np.sin(T) * (np.pi / n_theta) * (2 * np.pi / n_phi)
    c = np.clip(L @ np.asarray(normal, np.float64), 0, None)
    p_b = c / np.pi
    p_l = np.abs(L[..., 2]) / np.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        nee = np.where((L[..., 2] > 0) & (c > 0), (c / np.pi) * p_l / (p_l + p_b), 0.0)
        bsdf = np.where(c > 0, p_b * p_b / (p_b + p_l), 0.0)
    return float((nee * dw).sum()), float((bsdf * dw).sum())


def test_furnace_matches_closed_form():
    """reference furnace scene (whitefurnance.py FULLSCENESTR geometry: unit sphere seen by an orthographic camera along +z)
    with radiance 1 / albedo 1: the converged Y of a pixel must equal the closed form above for the pixel's normal."""
    size = 24
    src = WHITEFURNACE_FULL.format(hero="true", size=size).replace('(illuminant "D65")', "1").replace('"white"', "1").replace(":sample_count 8", ":sample_count 512")
    scene = prb.Scene.from_string(src)
    ora = OracleScene(scene)
    spp = 512
    r = ora.render(scene.tiles(4, 4), 0, spp, rng=scene.rng_map(), aov=False)
    Y = r["film"][..., 1]
    checked = 0
    for (px, py) in [(12, 12), (6, 12), (18, 12), (12, 5), (12, 19), (7, 7), (17, 16)]:
        # pixel centre -> point on the sphere (camera looks along +z from z = -1.00005, image plane [-1,1]^2, y up = -row)
        x = 2 * ((px + 0.5) / size - 0.5)
        y = -2 * ((py + 0.5) / size - 0.5)
        assert x * x + y * y < 0.8
        n = np.array([x, y, -np.sqrt(1 - x * x - y * y)])
        nee, bsdf = furnace_expectation(n)
        # the pixel averages over its footprint; the expectation is smooth, so the centre value is within the MC noise bar
        assert abs(Y[py, px] - (nee + bsdf)) < 0.04, ((px, py), Y[py, px], nee, bsdf)
        checked += 1
    assert checked == 7
    # the centre of the disc sees N = -z: no NEE contribution at all (the environment only samples its upper hemisphere), Y = 1/2
    nee0, bsdf0 = furnace_expectation([0, 0, -1])
    assert nee0 == 0.0 and abs(bsdf0 - 0.5) < 1e-3
    # straight up both strategies have the same pdf and every sample weighs 1/2 + 1/2
    nee1, bsdf1 = furnace_expectation([0, 0, 1])
    assert abs(nee1 + bsdf1 - 1.0) < 1e-3
