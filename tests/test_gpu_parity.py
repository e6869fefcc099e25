"""GPU parity tests proper (run on the B200 box: `pytest -m gpu`).  Everything goes through the C ABI
(include/prb200_abi.h); the oracle and the frozen golden fixtures are the checkers.

Bars (north star): hit entity / primitive ids (and u, v, t) BIT EXACT; same-seed film within a stated relative
RMSE with the IDENTICAL random-number consumption per pixel (RNG states after the render are compared), which
proves every path took the same decisions as the reference order of operations."""
import ctypes as C

import numpy as np
import pytest

import pearray_b200 as prb
from conftest import scene_path
from oracle_binding import OracleScene
from scene_strings import FURNACE, LPE_EXPRESSIONS, LPE_ZOO, MATERIAL_ZOO, MATERIAL_ZOO2, SKYSUN_ZOO
from test_golden_oracle import CBOX_CHANNEL_TOL, CBOX_LUMINANCE_TOL, GOLDEN, STAT_NAMES, cbox_reference_error, load_golden, load_scene

pytestmark = pytest.mark.gpu

ZOO2_WITH_SAMPLER = MATERIAL_ZOO2.replace("(sampler :slot 'aa' :type 'mjitt' :sample_count 16)",
                                          "(sampler :slot 'aa' :type '{0}' :sample_count 36) (sampler :slot 'lens' :type '{0}' :sample_count 36)")

# Same-seed renders are BIT EXACT against the oracle on every scene: all fp32 arithmetic on the device rounds like the
# oracle's (-fmad=false, IEEE div/sqrt, FTZ, explicit fma only where the reference has std::fma, correctly rounded
# transcendental functions on both sides, no reciprocal-multiplication of constant divisors -- tools/check_ptx_div.sh).
# The RNG state of every pixel after the render is compared too: it proves every path took the same decisions.
FILM_TOL = 0.0
RNG_EQUAL_MIN = 1.0


def make_ctx(scene, shading=None):
    ctx = prb.Context(0)
    ctx.upload_scene(scene)
    if shading is not None:
        ctx.set_shading_mode(shading)
    ctx.upload_rng(scene.rng_map())
    return ctx


def rel_rmse(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)) / max(1e-12, np.mean(np.abs(b))))


@pytest.mark.parametrize("name", GOLDEN)
def test_hits_bit_exact_vs_golden(name):
    g = load_golden(name)
    ctx = make_ctx(load_scene(name))
    for tag in ("cam", "inc"):
        ent, prim, u, v, t = ctx.trace_closest(g[tag + "_o"], g[tag + "_d"])
        assert np.array_equal(ent, g[tag + "_ent"]), "entity ids"
        assert np.array_equal(prim, g[tag + "_prim"]), "primitive ids"
        hit = ent != prb.INVALID_ID
        for a, b in ((u, g[tag + "_u"]), (v, g[tag + "_v"]), (t, g[tag + "_t"])):
            assert np.array_equal(a[hit].view(np.uint32), b[hit].view(np.uint32))
    occ = ctx.trace_any(g["inc_o"], g["inc_d"], None, g["inc_tmax"])
    assert np.array_equal(occ, g["inc_occ"])


# shading: -1 = the context times the single k_shade against the staged per-material-type kernels and alternates between them
# while it does (the default), 0 / 1 = pinned to one of them; the film must not depend on it
@pytest.mark.parametrize("shading", [-1, 0, 1])
@pytest.mark.parametrize("name", GOLDEN)
def test_film_vs_golden(name, shading):
    g = load_golden(name)
    scene = load_scene(name)
    if int(scene.desc.contents.n_lpe) and shading != 1:
        pytest.skip("scenes with light path expression channels shade staged")
    ctx = make_ctx(scene, shading)
    ctx.reset_stats()
    sx, sy, ex, ey = (int(x) for x in g["tile"])
    ctx.render_tiles([(sx, sy, ex, ey)], 0, 4)
    xyz, cnt = ctx.film()
    assert np.array_equal(cnt[sy:ey, sx:ex], g["count"])
    rng = ctx.download_rng().reshape(scene.height, scene.width)[sy:ey, sx:ex]
    eq = float(np.mean(rng == g["rng_after"]))
    assert eq >= RNG_EQUAL_MIN, "pixels with identical random-number consumption: %.5f" % eq
    # unfiltered film: undo nothing -- download applies the pixel filter, so compare against the filtered golden
    full = np.zeros((scene.height, scene.width, 3), np.float32)
    full[sy:ey, sx:ex] = g["film"]
    import oracle_binding as ob
    filt = np.empty_like(full)
    ob.lib().orc_apply_filter(scene.desc, full.ctypes.data_as(C.c_void_p), filt.ctypes.data_as(C.c_void_p))
    r = rel_rmse(xyz[sy:ey, sx:ex], filt[sy:ey, sx:ex])
    print(name, "relRMSE", r, "rng equal", eq)
    assert r <= FILM_TOL
    assert np.array_equal(xyz[sy:ey, sx:ex].view(np.uint32), filt[sy:ey, sx:ex].view(np.uint32)), "film bits"
    st = ctx.stats()
    assert [int(getattr(st, k)) for k in STAT_NAMES] == [int(v) for v in g["stats"]]


def test_depth_of_field_camera_rays_and_film_bit_exact():
    """PerspectiveCamera<HasDOF = true> (perspective.cpp:66-75): the lens sample moves the ray origin inside the aperture"""
    from scene_strings import DOF_ZOO
    scene = prb.Scene.from_string(DOF_ZOO.format(dof=":fstop 3 :aperture_radius 0.08"))
    assert scene.desc.contents.camera.has_dof == 1
    ctx = make_ctx(scene)
    ora = OracleScene(scene)
    tiles = [(0, 0, scene.width, scene.height)]
    for it in (0, 5):
        org, dr, wvl, pix = ctx.generate_camera_rays(tiles, it)
        oorg, odr, owvl, opix = ora.generate_camera_rays(tiles, it)
        assert np.array_equal(org.view(np.uint32), oorg.view(np.uint32)) and np.array_equal(dr.view(np.uint32), odr.view(np.uint32))
        assert np.ptp(org, axis=0).max() > 0.05  # the origins really spread over the aperture
    ctx.render_tiles(tiles, 0, 4)
    ref = ora.render(tiles, 0, 4)
    assert np.array_equal(ctx.film()[0].view(np.uint32), ref["filtered"].view(np.uint32))
    assert np.array_equal(ctx.download_rng(), ref["rng"])


def test_camera_rays_bit_exact():
    scene = load_scene("c2_cornellbox")
    ctx = make_ctx(scene)
    ora = OracleScene(scene)
    tiles = [(0, 0, 500, 500)]
    for it in (0, 3, 1023):
        org, dr, wvl, pix = ctx.generate_camera_rays(tiles, it)
        oorg, odr, owvl, opix = ora.generate_camera_rays(tiles, it)
        assert np.array_equal(pix, opix)
        assert np.array_equal(org.view(np.uint32), oorg.view(np.uint32))
        assert np.array_equal(dr.view(np.uint32), odr.view(np.uint32))
        assert np.array_equal(wvl.view(np.uint32), owvl.view(np.uint32))


@pytest.mark.parametrize("name", ["c1_sphere", "c4_boltsandgears"])
def test_sobol_and_mjitt_camera_rays_bit_exact(name):
    scene = load_scene(name)
    ctx = make_ctx(scene)
    ora = OracleScene(scene)
    tiles = [(100, 100, 228, 228)]
    for it in (0, 7):
        a = ctx.generate_camera_rays(tiles, it)
        b = ora.generate_camera_rays(tiles, it)
        for x, y in zip(a, b):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


@pytest.mark.parametrize("sampler", ["stratified", "uniform", "halton", "hammersley"])
def test_stratified_and_uniform_samplers_bit_exact(sampler):
    """StratifiedSampler.cpp / UniformSampler.cpp / HaltonSampler.cpp: camera rays and a short render against the oracle"""
    scene = prb.Scene.from_string(ZOO2_WITH_SAMPLER.format(sampler))
    ctx = make_ctx(scene)
    ora = OracleScene(scene)
    tiles = [(0, 0, 32, 32)]
    for it in (0, 5, 40):  # 40 >= sample_count: indices beyond the strata / the precomputed table
        for x, y in zip(ctx.generate_camera_rays(tiles, it), ora.generate_camera_rays(tiles, it)):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
    ctx.render_tiles(tiles, 0, 4)
    xyz, cnt = ctx.film()
    ref = ora.render(tiles, 0, 4)
    assert np.array_equal(xyz.view(np.uint32), ref["filtered"].view(np.uint32))
    assert np.array_equal(ctx.download_rng(), ref["rng"])


@pytest.mark.parametrize("mapper,kind", [("'cie'", 3), ("'cie_y'", 3), ("'agh'", 4), ("'agh' :cmis false", 5)])
def test_cie_and_agh_spectral_mappers_bit_exact(mapper, kind):
    """spectralmapper/cie.cpp (four independent wavelengths from the CIE X+Y+Z / Y CDF) and agh.cpp (lambda = B - atanh(C - N u) / A,
    CMIS and hero form; its pdf 1 / (cosh^2 N) is restated as the reference writes it) -- SURVEY 8(f)-3"""
    src = MATERIAL_ZOO2.replace("(sampler :slot 'aa'", "(spectral_mapper :slot 'pixel' :type %s) (sampler :slot 'aa'" % mapper)
    scene = prb.Scene.from_string(src)
    assert scene.desc.contents.pixel_mapper.type == kind
    ctx = make_ctx(scene)
    ora = OracleScene(scene)
    tiles = [(0, 0, 32, 32)]
    for it in (0, 9):
        a, b = ctx.generate_camera_rays(tiles, it), ora.generate_camera_rays(tiles, it)
        for x, y in zip(a, b):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
        wvl = a[2]
        assert wvl.min() >= 390.0 and wvl.max() <= 830.0
    ctx.render_tiles(tiles, 0, 4)
    xyz, cnt = ctx.film()
    ref = ora.render(tiles, 0, 4)
    assert np.array_equal(xyz.view(np.uint32), ref["filtered"].view(np.uint32))
    assert np.array_equal(ctx.download_rng(), ref["rng"])


@pytest.mark.parametrize("zoo", ["zoo", "zoo4"])
def test_material_unit_calls_vs_oracle(zoo):
    """IMaterial::eval / ::sample through prb_material_eval / prb_material_sample for every material of the zoo (zoo4: blend / add
    materials, nested up to three levels)"""
    from scene_strings import MATERIAL_ZOO4
    scene = prb.Scene.from_string(MATERIAL_ZOO if zoo == "zoo" else MATERIAL_ZOO4)
    ctx = make_ctx(scene)
    ora = OracleScene(scene)
    rs = np.random.RandomState(7)
    n = 512
    for mat in range(scene.desc.contents.n_materials):
        q = (prb.MaterialQuery * n)()
        for i in range(n):
            v = rs.normal(size=3); v /= np.linalg.norm(v)
            l = rs.normal(size=3); l /= np.linalg.norm(l)
            q[i].V[:] = [float(x) for x in v]
            q[i].L[:] = [float(x) for x in l]
            q[i].wavelength_nm[:] = [float(x) for x in rs.uniform(400, 780, 4)]
            q[i].uv[:] = [float(x) for x in rs.uniform(0, 1, 2)]
            q[i].ray_flags = 1
            q[i].material_id = mat
            q[i].rng_state = int(rs.randint(1, 2 ** 62)) | 3
        for kind in ("eval", "sample"):
            g = getattr(ctx, "material_" + kind)(q)
            o = getattr(ora, "material_" + kind)(q)
            ga = np.array([[*r.weight, *r.pdf_s, *r.L] for r in g], dtype=np.float32)
            oa = np.array([[*r.weight, *r.pdf_s, *r.L] for r in o], dtype=np.float32)
            gf = np.array([(r.flags, r.type, r.rng_state) for r in g], dtype=np.uint64)
            of = np.array([(r.flags, r.type, r.rng_state) for r in o], dtype=np.uint64)
            assert np.array_equal(gf, of), (mat, kind, "flags / scattering type / rng state")
            same = (ga.view(np.uint32) == oa.view(np.uint32)) | (np.isnan(ga) & np.isnan(oa))
            assert same.all(), (mat, kind, "weight / pdf / direction bits", int((~same).sum()))


@pytest.mark.parametrize("interp", ["closest", "bilinear", "bicubic"])
def test_textured_scene_bit_exact_vs_oracle(tmp_path, interp):
    """an image texture as albedo (NonParametricImageNode: texel fetch, sRGB linearisation, RGB ->
    spectrum coefficient-cube lookup per shading point) under an image-based environment light (Distribution2D sampling)"""
    from test_textures import SCENE, write_pfm, write_ppm
    rs = np.random.RandomState(11)
    write_ppm(str(tmp_path / "albedo.ppm"), rs.randint(0, 256, size=(9, 13, 3)).astype(np.uint8))  # sRGB encoded
    env = np.full((8, 16, 3), 0.05, np.float32)
    env[2:4, 5:7] = (0.9, 0.8, 0.7)
    write_pfm(str(tmp_path / "env.pfm"), env)
    src = SCENE.replace("(light :type 'env' :radiance %(env)s)", "(texture :name 'sky' :type 'color' :file '%s' :interpolation '%s' :wrap 'periodic')\n"
                        " (light :type 'env' :radiance (texture 'sky'))" % (tmp_path / "env.pfm", interp))
    scene = prb.Scene.from_string(src % dict(file=tmp_path / "albedo.ppm", options=":interpolation '%s' :wrap 'mirror'" % interp, env="1"))
    d = scene.desc.contents
    assert sum(d.nodes[i].type == 7 for i in range(d.n_nodes)) == 2 and d.lights[0].dist_w == 16
    ctx = make_ctx(scene)
    ora = OracleScene(scene)
    tiles = [(0, 0, scene.width, scene.height)]
    ctx.render_tiles(tiles, 0, 6)
    xyz, cnt = ctx.film()
    ref = ora.render(tiles, 0, 6)
    assert np.array_equal(ctx.download_rng(), ref["rng"])
    assert np.array_equal(xyz.view(np.uint32), ref["filtered"].view(np.uint32)) and xyz.max() > 0
    # IMaterial::eval at random surface parameters, also outside [0, 1] (wrap mode)
    q = (prb.MaterialQuery * 256)()
    for i in range(256):
        v = rs.normal(size=3); v[2] = abs(v[2]); v /= np.linalg.norm(v)
        l = rs.normal(size=3); l[2] = abs(l[2]); l /= np.linalg.norm(l)
        q[i].V[:] = [float(x) for x in v]
        q[i].L[:] = [float(x) for x in l]
        q[i].wavelength_nm[:] = [float(x) for x in rs.uniform(400, 780, 4)]
        q[i].uv[:] = [float(x) for x in rs.uniform(-0.5, 1.5, 2)]
        q[i].ray_flags = 1
        q[i].material_id = 0
        q[i].rng_state = 3
    g, o = ctx.material_eval(q), ora.material_eval(q)
    ga = np.array([[*r.weight, *r.pdf_s] for r in g], dtype=np.float32)
    oa = np.array([[*r.weight, *r.pdf_s] for r in o], dtype=np.float32)
    assert np.array_equal(ga.view(np.uint32), oa.view(np.uint32)) and ga[:, :4].max() > 0


def test_soup_hits_bit_exact_vs_oracle():
    """synthetic triangle soup (SURVEY 8(d) C5 at a size the oracle finishes in seconds): primary, shadow and incoherent
    bounce rays"""
    scene = prb.Scene.soup(200000, seed=1234, film=(256, 256))
    ctx = make_ctx(scene)
    ora = OracleScene(scene)
    org, dr, wvl, pix = ctx.generate_camera_rays([(0, 0, 256, 256)], 0)
    got = ctx.trace_closest(org, dr)
    ref = ora.trace_closest(org, dr)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    hit = got[0] != prb.INVALID_ID
    assert hit.mean() > 0.1
    for a, b in zip(got[2:], ref[2:]):
        assert np.array_equal(a[hit].view(np.uint32), b[hit].view(np.uint32))
    P = (org[hit] + dr[hit] * got[4][hit, None]).astype(np.float32)
    light = np.array([0, 3, 0], np.float32)
    d = light - P
    dist = np.linalg.norm(d, axis=1).astype(np.float32)
    d = (d / dist[:, None]).astype(np.float32)
    tmin = np.full(len(P), 1e-4, np.float32)
    tmax = (dist - 1e-3).astype(np.float32)
    assert np.array_equal(ctx.trace_any(P, d, tmin, tmax), ora.trace_any(P, d, tmin, tmax))
    rs = np.random.RandomState(5)
    b = rs.normal(size=P.shape)
    b = (b / np.linalg.norm(b, axis=1, keepdims=True)).astype(np.float32)
    g2 = ctx.trace_closest(P, b)
    o2 = ora.trace_closest(P, b)
    assert np.array_equal(g2[0], o2[0]) and np.array_equal(g2[1], o2[1])
    h2 = g2[0] != prb.INVALID_ID
    assert np.array_equal(g2[4][h2].view(np.uint32), o2[4][h2].view(np.uint32))


def test_soup_large_properties():
    """size-independent properties on a soup far too large for the oracle to trace in test time (2 M triangles):
    any-hit == (closest-hit exists inside the interval); the closest hit is stable under shortening tmax to just behind it;
    the reported t reproduces the hit when the ray is restarted past it."""
    scene = prb.Scene.soup(2000000, seed=7, film=(1024, 1024))
    ctx = make_ctx(scene)
    org, dr, wvl, pix = ctx.generate_camera_rays([(0, 0, 1024, 1024)], 0)
    ent, prim, u, v, t = ctx.trace_closest(org, dr)
    hit = ent != prb.INVALID_ID
    assert 0.2 < hit.mean() <= 1.0
    tmax = np.where(hit, t * np.float32(1.5), np.float32(10.0)).astype(np.float32)
    occ = ctx.trace_any(org, dr, None, tmax)
    assert np.array_equal(occ.astype(bool), hit)
    # nothing is hit strictly before the closest hit
    before = np.where(hit, np.nextafter(t, np.float32(0)), np.float32(10.0)).astype(np.float32)
    occ2 = ctx.trace_any(org, dr, None, before)
    ent2, prim2, _, _, t2 = ctx.trace_closest(org, dr, None, t.copy())
    assert np.array_equal(ent2[hit], ent[hit]) and np.array_equal(prim2[hit], prim[hit]) and np.array_equal(t2[hit], t[hit])
    # occ2 may only be set where ANOTHER primitive lies within the same float t (ties); must be rare
    assert occ2[hit].mean() < 1e-3
    assert (u[hit] >= 0).all() and (v[hit] >= 0).all() and (u[hit] + v[hit] <= 1 + 1e-5).all()


def test_tile_partition_bit_identical_on_gpu():
    """interleaved tile ownership -> the summed per-rank films are bit-identical to the 1-GPU film.
    Two contexts on the same device play the two ranks; films are combined with prb_film_export_device buffers."""
    import torch
    from pearray_b200 import multigpu
    scene = load_scene("c3_cornellbox_glassy")
    tiles = scene.tiles(8, 8)
    spp = 4
    single = make_ctx(scene)
    single.render_tiles(tiles, 0, spp)
    ref = torch.zeros(scene.height * scene.width * 4, dtype=torch.float32, device="cuda")
    single.film_export_device(ref.data_ptr())
    acc = torch.zeros_like(ref)
    for rank in range(2):
        c = make_ctx(scene)
        c.render_tiles(multigpu.partition_tiles(tiles, rank, 2), 0, spp)
        buf = torch.zeros_like(ref)
        c.film_export_device(buf.data_ptr())
        acc += buf
    torch.cuda.synchronize()
    assert torch.equal(acc.view(torch.int32), ref.view(torch.int32))
    # import the reduced film back and download through the filter: equals the single-context download
    single2 = make_ctx(scene)
    single2.film_import_device(acc.data_ptr())
    a, ca = single2.film()
    b, cb = single.film()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(ca, cb)


def test_resume_and_rerun_determinism():
    scene = load_scene("c2_cornellbox")
    tile = [(200, 200, 264, 264)]
    a = make_ctx(scene)
    a.render_tiles(tile, 0, 6)
    fa, ca = a.film()
    b = make_ctx(scene)
    b.render_tiles(tile, 0, 2)
    b.render_tiles(tile, 2, 4)
    fb, cb = b.film()
    assert np.array_equal(fa.view(np.uint32), fb.view(np.uint32)) and np.array_equal(ca, cb)
    assert np.array_equal(a.download_rng(), b.download_rng())


def test_full_frame_c2_vs_oracle_and_stats():
    """whole 500x500 film, 2 iterations: every pixel, all 64 tiles"""
    scene = load_scene("c2_cornellbox")
    ctx = make_ctx(scene)
    ctx.reset_stats()
    tiles = scene.tiles(8, 8)
    ctx.render_tiles(tiles, 0, 2)
    xyz, cnt = ctx.film()
    ref = OracleScene(scene).render(tiles, 0, 2)
    assert np.array_equal(cnt, ref["count"])
    # identical decisions everywhere: same random-number consumption of all 250 000 pixels, same counters
    assert np.array_equal(ctx.download_rng(), ref["rng"])
    st = ctx.stats()
    assert {k: int(getattr(st, k)) for k in STAT_NAMES} == ref["stats"]
    r = rel_rmse(xyz, ref["filtered"])
    print("full frame relRMSE", r)
    assert np.array_equal(xyz.view(np.uint32), ref["filtered"].view(np.uint32))
    aov = ctx.film_aov()
    assert np.array_equal(aov.view(np.uint32), ref["aov"].view(np.uint32))
    with pytest.raises(prb.PrbError):  # opt in: prb_settings.want_aov_ext (the host sets it from the scene's output channels)
        ctx.film_aov_ext()
    scene.settings.want_aov_ext = 1
    ctx2 = make_ctx(scene)
    ctx2.render_tiles(tiles, 0, 2)
    ext = ctx2.film_aov_ext()  # tangent, bitangent, view, material id, emission id (LocalFrameOutputDevice.cpp:268-284)
    assert np.array_equal(ext.view(np.uint32), ref["aov_ext"].view(np.uint32)) and np.abs(ext[..., :9]).max() > 0
    assert np.array_equal(ctx2.film()[0].view(np.uint32), xyz.view(np.uint32))


@pytest.mark.parametrize("name,iters", [("c4_boltsandgears", 2), ("c4c_complex", 1)])
def test_full_frame_large_films_vs_oracle(name, iters):
    """whole 1000x1000 / 1920x1080 films: the only sizes at which k_shade sorts multi-pass windows (3 and 4 passes of 512
    slots per block) and, for complex.prc, the persistent k_trace runs over every pixel; same bits, RNG states and counters
    as the oracle"""
    scene = load_scene(name)
    ctx = make_ctx(scene)
    ctx.reset_stats()
    tiles = scene.tiles(8, 8)
    ctx.render_tiles(tiles, 0, iters)
    xyz, cnt = ctx.film()
    ref = OracleScene(scene).render(tiles, 0, iters)
    assert np.array_equal(cnt, ref["count"])
    assert np.array_equal(ctx.download_rng(), ref["rng"])
    st = ctx.stats()
    assert {k: int(getattr(st, k)) for k in STAT_NAMES} == ref["stats"]
    assert np.array_equal(xyz.view(np.uint32), ref["filtered"].view(np.uint32))


def test_shading_mode_is_measured_and_does_not_change_the_film():
    """a mixed-material scene starts undecided, decides within the first 4 poll intervals (128 wavefront iterations) and
    renders the same bits as either pinned mode"""
    scene = load_scene("c4_boltsandgears")
    tiles = [(200, 200, 456, 328)]
    films = {}
    for mode in (-1, 0, 1):
        ctx = make_ctx(scene, mode)
        assert ctx.shading_mode() == mode
        ctx.render_tiles(tiles, 0, 48)
        films[mode] = (ctx.film()[0], ctx.download_rng(), ctx.shading_mode())
        ctx.close()
    assert films[-1][2] in (0, 1), "still undecided after a 48-spp render"
    for mode in (0, 1):
        assert np.array_equal(films[mode][0].view(np.uint32), films[-1][0].view(np.uint32))
        assert np.array_equal(films[mode][1], films[-1][1])
    lam = make_ctx(load_scene("c2_cornellbox"))
    assert lam.shading_mode() == 0, "an all-Lambert scene always takes the inlined single kernel"


def test_lpe_channels_bit_exact_vs_oracle(tmp_path):
    """one spectral channel per light path expression; the device carries the DFA state of every expression per
    slot, the oracle re-walks the token string of every fragment from the start (LightPathExpression::match)"""
    scene = prb.Scene.from_string(LPE_ZOO)
    ctx = make_ctx(scene)
    assert ctx.shading_mode() == 1, "scenes with LPE channels shade staged"
    tiles = [(0, 0, scene.width, scene.height)]
    ref = OracleScene(scene).render(tiles, 0, 6)
    ctx.render_tiles(tiles, 0, 4)
    ctx.render_tiles(tiles, 4, 2)  # the channels resume like the main film
    xyz, cnt = ctx.film()
    assert np.array_equal(xyz.view(np.uint32), ref["filtered"].view(np.uint32))
    assert np.array_equal(ctx.download_rng(), ref["rng"])
    for k, expr in enumerate(LPE_EXPRESSIONS):
        ch = ctx.film_lpe(k)
        assert np.array_equal(ch.view(np.uint32), ref["lpe_filtered"][k].view(np.uint32)), expr
        assert ch.max() > 0, expr
    assert np.array_equal(ctx.film_lpe(LPE_EXPRESSIONS.index("C.*L")).view(np.uint32), xyz.view(np.uint32))
    with pytest.raises(prb.PrbError):
        ctx.film_lpe(len(LPE_EXPRESSIONS))
    with pytest.raises(prb.PrbError):
        ctx.set_shading_mode(0)
    # a re-render from iteration 0 starts the channels from zero again
    ctx.upload_rng(scene.rng_map())
    ctx.render_tiles(tiles, 0, 6)
    assert np.array_equal(ctx.film_lpe(1).view(np.uint32), ref["lpe_filtered"][1].view(np.uint32))


def test_reference_plane_and_sphere_known_answers_on_device():
    """the Plane / Sphere intersection cases of the reference's own tests (src/tests/plane.cpp:66-117, sphere.cpp:77-110) through
    prb_trace_closest: the device returns what the oracle returns, and the oracle is checked against the reference's expected
    values in test_oracle_known_answers.py"""
    from test_oracle_known_answers import GEOMETRY_SCENE
    cases = [("(entity :name 'ball' :type 'sphere' :radius 1 :material 'white')", [[-2, 0, 0], [-2, 0, 0], [0, 0, 0]], [[1, 0, 0], [-1, 0, 0], [1, 0, 0]]),
             ("(entity :name 'quad' :type 'plane' :x_axis [1,0,0] :y_axis [0,1,0] :material 'white')", [[0.5, 0.5, -1], [0.5, 0.5, -1]], [[0, 0, 1], [0, 1, 0]]),
             ("(entity :name 'quad' :type 'plane' :x_axis [10,0,0] :y_axis [0,20,0] :material 'white')", [[5, 10, -1]], [[0, 0, 1]])]
    for ent_src, org, dr in cases:
        scene = prb.Scene.from_string(GEOMETRY_SCENE % ent_src)
        ctx = make_ctx(scene)
        org, dr = np.array(org, np.float32), np.array(dr, np.float32)
        tmin = np.zeros(len(org), np.float32)
        got = ctx.trace_closest(org, dr, tmin)
        ref = OracleScene(scene).trace_closest(org, dr, tmin=tmin)
        hit = ref[0] != 0xFFFFFFFF
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
        for k in (2, 3, 4):
            assert np.array_equal(got[k][hit].view(np.uint32), ref[k][hit].view(np.uint32))
        assert hit[0] and abs(float(got[4][0]) - 1) <= 2 * np.finfo(np.float32).eps


def _many_faces_scene(seed, n_meshes=3, tris_per_mesh=14):
    """a small scene (<= 64 triangles, <= 16 entities: k_trace_small) whose triangles do NOT pair into quads, so that it has more
    than 32 faces (second candidate word of traverseSmall), in rotated / scaled instances (world-space face boxes)"""
    rs = np.random.RandomState(seed)
    fmt = lambda a: ",".join("[%.6f, %.6f, %.6f]" % tuple(v) for v in a)
    parts = ["""(scene :render_width 48 :render_height 48 :camera 'Camera'
(integrator :type 'direct' :max_ray_depth 5)
(sampler :slot 'aa' :type 'mjitt' :sample_count 16)
(camera :name 'Camera' :type 'standard' :width 0.72 :height 0.72 :local_direction [0,0,-1] :local_up [0,1,0] :local_right [1,0,0]
 :near 0.1 :far 100.0 :transform [1.0,0.0,0.0,0.0,0.0,0.0,-1.0,-3.9,0.0,1.0,0.0,1.0,0.0,0.0,0.0,1.0])
(emission :name 'light_em' :type 'standard' :radiance (smul (illuminant "D65") (illum 17 12 4)))
(material :name 'light' :type 'diffuse' :albedo (refl 0.78 0.78 0.78))
(material :name 'floor' :type 'diffuse' :albedo (refl 0.725 0.71 0.68))
(material :name 'red' :type 'diffuse' :albedo (refl 0.63 0.065 0.05))
(material :name 'green' :type 'diffuse' :albedo (refl 0.14 0.45 0.091))
(mesh :name 'lightm' (attribute :type 'p' [-0.4,-0.3,1.98],[-0.4,0.3,1.98],[0.4,0.3,1.98],[0.4,-0.3,1.98])
 (attribute :type 'n' [0,0,-1],[0,0,-1],[0,0,-1],[0,0,-1]) (faces [0,1,2],[0,2,3]))
(entity :name 'lightm' :type 'mesh' :materials 'light' :emission 'light_em' :mesh 'lightm')
(mesh :name 'floorm' (attribute :type 'p' [-1.5,-1.5,0],[1.5,-1.5,0],[1.5,1.5,0],[-1.5,1.5,0])
 (attribute :type 'n' [0,0,1],[0,0,1],[0,0,1],[0,0,1]) (faces [0,1,2],[0,2,3]))
(entity :name 'floorm' :type 'mesh' :materials 'floor' :mesh 'floorm')
"""]
    for m in range(n_meshes):
        centres = rs.uniform([-0.8, -0.8, 0.1], [0.8, 0.8, 1.6], size=(tris_per_mesh, 3))
        verts = (centres[:, None, :] + rs.normal(scale=0.22, size=(tris_per_mesh, 3, 3))).reshape(-1, 3)
        nrm = np.repeat(np.cross(verts[1::3] - verts[0::3], verts[2::3] - verts[0::3]), 3, axis=0)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        faces = ",".join("[%d,%d,%d]" % (3 * i, 3 * i + 1, 3 * i + 2) for i in range(tris_per_mesh))
        a = rs.uniform(0, 2 * np.pi)
        sc = rs.uniform(0.7, 1.2)
        c, s_ = np.cos(a) * sc, np.sin(a) * sc
        xf = [c, -s_, 0.0, rs.uniform(-0.2, 0.2), s_, c, 0.0, rs.uniform(-0.2, 0.2), 0.0, 0.0, sc, rs.uniform(0.0, 0.2), 0.0, 0.0, 0.0, 1.0]
        parts.append("(mesh :name 'soup%d' (attribute :type 'p' %s) (attribute :type 'n' %s) (faces %s))\n" % (m, fmt(verts), fmt(nrm), faces))
        parts.append("(entity :name 'soup%d' :type 'mesh' :materials '%s' :mesh 'soup%d' :transform [%s])\n"
                     % (m, ("red", "green", "floor")[m % 3], m, ",".join("%.6f" % v for v in xf)))
    parts.append(")")
    return "".join(parts)


@pytest.mark.parametrize("seed", [1, 2])
def test_small_scene_with_more_than_32_faces_vs_oracle(seed):
    """k_trace_small past its first candidate word: 44 faces (random triangles do not pair into quads) in rotated, scaled
    instances; films, per-pixel RNG states and the statistics bit exact against the oracle (whose triangle loop for meshes of
    <= 16 triangles is exhaustive: the box phase of the device must not cull anything the watertight test accepts)"""
    scene = prb.Scene.from_string(_many_faces_scene(seed))
    d = scene.desc.contents
    assert 32 < d.n_bvh_tris <= 64 and d.n_entities <= 16
    ctx = make_ctx(scene)
    tiles = [(0, 0, scene.width, scene.height)]
    ctx.render_tiles(tiles, 0, 8)
    xyz, cnt = ctx.film()
    ref = OracleScene(scene).render(tiles, 0, 8)
    assert np.array_equal(ctx.download_rng(), ref["rng"])
    assert np.array_equal(xyz.view(np.uint32), ref["filtered"].view(np.uint32)) and xyz.max() > 0
    st = ctx.stats()
    for k in STAT_NAMES:
        assert int(getattr(st, k)) == int(ref["stats"][k]), k


def test_uniform_non_lambert_scene_uses_the_generic_single_pass_kernel():
    """all materials of one non-Lambert type: k_shade<128, 1, leaf dispatch> (the Cornell box takes the all-Lambert
    instantiation, mixed scenes the 512-thread one)"""
    src = MATERIAL_ZOO2
    for m in ("'mirror'", "'reflection' :specularity (refl 0.9 0.6 0.2)", "'diffuse' :albedo 0.5"):
        src = src.replace(":type %s" % m, ":type 'orennayar' :albedo (refl 0.6 0.5 0.4) :roughness 0.4")
    src = src.replace(":type 'rough' ", ":type 'orennayar' ")
    scene = prb.Scene.from_string(src)
    d = scene.desc.contents
    assert {d.materials[i].type for i in range(d.n_materials)} == {7}
    ctx = make_ctx(scene)
    tiles = [(0, 0, 32, 32)]
    ctx.render_tiles(tiles, 0, 4)
    xyz, _ = ctx.film()
    ref = OracleScene(scene).render(tiles, 0, 4)
    assert np.array_equal(xyz.view(np.uint32), ref["filtered"].view(np.uint32))
    assert np.array_equal(ctx.download_rng(), ref["rng"])


def test_furnace_on_gpu():
    scene = prb.Scene.from_string(FURNACE % dict(hero="true"))
    ctx = make_ctx(scene)
    ctx.render_tiles([(0, 0, 48, 48)], 0, 64)
    f, cnt = ctx.film()
    ys, xs = np.mgrid[0:48, 0:48]
    rad = np.hypot(xs - 23.5, ys - 23.5)
    assert abs(f[rad < 12][:, 1].mean() - 1.0) < 0.03
    assert abs(f[rad > 22][:, 1].mean() - 4.0) < 0.12


def test_stage_profiling_entry_points():
    scene = load_scene("c2_cornellbox")
    ctx = make_ctx(scene)
    tile = [(0, 0, 128, 128)]
    ctx.render_tiles(tile, 0, 2)
    plain, _ = ctx.film()
    ctx2 = make_ctx(scene)
    ctx2.set_profiling(True)
    ctx2.render_tiles(tile, 0, 2)
    st = ctx2.stage_times()
    assert set(st) == {"trace", "shade"}
    assert all(ms > 0 and n > 0 for ms, n in st.values())
    prof, _ = ctx2.film()
    assert np.array_equal(plain.view(np.uint32), prof.view(np.uint32))  # profiling does not change results
    assert ctx2.last_device_ms() > 0


def test_host_render_context_matches_abi_path():
    """the C++ host driver (RenderContext::start -> IIntegratorInstance::onTile -> prb_render_tiles) produces the same
    film as driving the C ABI directly"""
    h = prb.host_lib()
    scene = load_scene("c3_cornellbox_glassy")
    rc = h.prh_render_context_create(scene._h, 0, 0, 1)
    assert rc, h.prh_last_error()
    assert h.prh_render_context_start(rc, 8, 8, 3) == 0
    h.prh_render_context_wait(rc)
    dev = h.prh_render_context_device(rc)
    a = np.empty((scene.height, scene.width, 3), np.float32)
    assert prb.device_lib().prb_film_download(dev, a.ctypes.data_as(C.c_void_p), None) == 0
    h.prh_render_context_destroy(rc)
    scene2 = load_scene("c3_cornellbox_glassy")
    ctx = make_ctx(scene2)
    ctx.render_tiles(scene2.tiles(8, 8), 0, 3)
    b, _ = ctx.film()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_host_render_context_writes_the_output_files(tmp_path):
    """RenderContext::saveOutputs: results/<name>.exr of the scene's (output ...) blocks holds the film the C ABI returns
    (reference OutputSpecification::save + ImageWriter::save)"""
    from test_image_output import SCENE, read_exr
    h = prb.host_lib()
    scene = prb.Scene.from_string(SCENE)
    rc = h.prh_render_context_create(scene._h, 0, 0, 1)
    assert rc, h.prh_last_error()
    assert h.prh_render_context_start(rc, 8, 8, 4) == 0
    h.prh_render_context_wait(rc)
    dev = h.prh_render_context_device(rc)
    xyz = np.empty((scene.height, scene.width, 3), np.float32)
    cnt = np.empty((scene.height, scene.width), np.uint32)
    assert prb.device_lib().prb_film_download(dev, xyz.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p)) == 0
    assert h.prh_render_context_save_outputs(rc, str(tmp_path).encode()) == 2
    h.prh_render_context_destroy(rc)
    _, _, pl = read_exr(str(tmp_path / "results" / "aovs.exr"))
    for k, c in enumerate("RGB"):  # :color 'xyz'
        assert np.array_equal(pl[c], xyz[..., k])
        assert np.array_equal(pl["[C.*L]." + c], xyz[..., 1])  # :color 'lum' of an expression that accepts every path: Y of the same film
    assert np.array_equal(pl["sample_count"], cnt.astype(np.float32))
    assert cnt.max() == 4 and xyz.max() > 0
    _, chans, _ = read_exr(str(tmp_path / "results" / "image.exr"))
    assert chans == ["B", "G", "R"]


def test_invalid_arguments_on_device():
    scene = load_scene("c2_cornellbox")
    ctx = make_ctx(scene)
    with pytest.raises(prb.PrbError):
        ctx.render_tiles([(0, 0, 501, 10)], 0, 1)  # tile outside the film
    with pytest.raises(prb.PrbError):
        ctx.upload_rng(np.zeros(10, np.uint64))
    fresh = prb.Context(0)
    with pytest.raises(prb.PrbError):
        fresh.render_tiles([(0, 0, 8, 8)], 0, 1)  # no scene
    # empty inputs are no-ops
    ctx.render_tiles([], 0, 1) if False else None
    e = ctx.trace_closest(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert len(e[0]) == 0


def test_gpu_vs_reference_golden_image():
    """full 128 spp render of the evaluation scene against the reference's golden image (cross-renderer sanity bound)"""
    scene = load_scene("c0_evaluation")
    ctx = make_ctx(scene)
    ctx.render_tiles(scene.tiles(8, 8), 0, 128)
    xyz, _ = ctx.film()
    lum_err, ratio = cbox_reference_error(xyz)
    assert lum_err < CBOX_LUMINANCE_TOL, lum_err
    assert np.all(np.abs(ratio - 1) < CBOX_CHANNEL_TOL), ratio


# ------------------------------------------------------------------ round 2: the reference's furnace test, monochrome modes, feedback
@pytest.mark.parametrize("mode", ["spec", "non_hero", "full"])
def test_whitefurnance_gpu(mode):
    """literal port of src/tests/python/whitefurnance.py on the device (tests/test_whitefurnance.py holds the derivation of the
    asserted values): 200x200, hammersley 8 spp, block filter radius 0, orthographic camera, ctx.start(8, 8); film, sample
    counts, feedback bits, RNG states and counters bit-equal to the oracle, and the reference test's probes on the device film"""
    import test_whitefurnance as wf
    scene = prb.Scene.from_string(wf.MODES[mode])
    ctx = make_ctx(scene)
    ctx.reset_stats()
    tiles = scene.tiles(8, 8)
    ctx.render_tiles(tiles, 0, 8)
    xyz, cnt = ctx.film()
    fb = ctx.film_feedback()
    ref = OracleScene(scene).render(tiles, 0, 8, rng=scene.rng_map())
    assert np.array_equal(cnt, ref["count"])
    assert np.array_equal(ctx.download_rng(), ref["rng"])
    assert np.array_equal(fb, ref["feedback"])
    st = ctx.stats()
    assert {k: int(getattr(st, k)) for k in STAT_NAMES} == ref["stats"]
    assert np.array_equal(xyz.view(np.uint32), ref["filtered"].view(np.uint32))
    if mode == "spec":
        wf.check_spec(xyz, fb, cnt)
    else:
        other = "non_hero" if mode == "full" else "full"
        s2 = prb.Scene.from_string(wf.MODES[other])
        c2 = make_ctx(s2)
        c2.render_tiles(s2.tiles(8, 8), 0, 8)
        x2, _ = c2.film()
        f2 = c2.film_feedback()
        if mode == "full":
            wf.check_cie_modes(xyz, x2, fb, f2, cnt)
        else:
            wf.check_cie_modes(x2, xyz, f2, fb, cnt)


@pytest.mark.parametrize("variant", ["non_hero", "single_wavelength", "single_wavelength_zoo", "non_hero_zoo"])
def test_monochrome_modes_vs_oracle(variant):
    """`:spectral_hero false` (forced-monochrome rays, ordinary film) and `:spectral_domain <nm>` (monotonic film) against the
    oracle: film, feedback bits, RNG states; on the furnace and on the material zoo (NEE fragments of monochrome rays are NaN
    as direct.cpp is written, delta materials and area lights included)"""
    if variant == "non_hero":
        src, tile, it = FURNACE % dict(hero="false"), (0, 0, 48, 48), 16
    elif variant == "single_wavelength":
        src, tile, it = (FURNACE % dict(hero="true")).replace(":spectral_hero true", ":spectral_hero true :spectral_domain 520"), (0, 0, 48, 48), 16
    elif variant == "single_wavelength_zoo":
        src, tile, it = SKYSUN_ZOO.replace(":camera 'Camera'", ":camera 'Camera' :spectral_domain 610"), (0, 0, 48, 48), 8
    else:
        src, tile, it = SKYSUN_ZOO.replace(":camera 'Camera'", ":camera 'Camera' :spectral_hero false"), (0, 0, 48, 48), 8
    scene = prb.Scene.from_string(src)
    assert bool(scene.settings.film_monotonic) == variant.startswith("single_wavelength")
    ctx = make_ctx(scene)
    ctx.render_tiles([tile], 0, it)
    xyz, cnt = ctx.film()
    ref = OracleScene(scene).render([tile], 0, it, rng=scene.rng_map())
    assert np.array_equal(ctx.download_rng(), ref["rng"]) and np.array_equal(cnt, ref["count"])
    assert np.array_equal(ctx.film_feedback(), ref["feedback"])
    assert ref["feedback"].any()
    assert np.array_equal(xyz.view(np.uint32), ref["filtered"].view(np.uint32))
    if variant.startswith("single_wavelength"):
        assert np.array_equal(xyz[..., 0], xyz[..., 1]) and np.array_equal(xyz[..., 0], xyz[..., 2])


def test_feedback_is_zero_on_the_config_scenes():
    for name in ("c2_cornellbox", "c3_cornellbox_glassy"):
        g = load_golden(name)
        scene = load_scene(name)
        ctx = make_ctx(scene)
        sx, sy, ex, ey = (int(x) for x in g["tile"])
        ctx.render_tiles([(sx, sy, ex, ey)], 0, 4)
        ref = OracleScene(scene).render([(sx, sy, ex, ey)], 0, 4, rng=scene.rng_map())
        assert np.array_equal(ctx.film_feedback(), ref["feedback"])


def test_render_tiles_validation_and_rerender():
    """overlapping tiles and a missing RNG map are rejected; a second render from iteration 0 starts the
    counters from scratch; the cached CUDA graph serves calls with different iteration ranges"""
    scene = load_scene("c2_cornellbox")
    fresh = prb.Context(0)
    fresh.upload_scene(scene)
    with pytest.raises(prb.PrbError):
        fresh.render_tiles([(0, 0, 8, 8)], 0, 1)  # prb_upload_rng not called
    ctx = make_ctx(scene)
    with pytest.raises(prb.PrbError):
        ctx.render_tiles([(0, 0, 16, 16), (8, 8, 24, 24)], 0, 1)  # overlap
    tile = [(100, 100, 164, 164)]
    ctx.render_tiles(tile, 0, 3)
    a, ca = ctx.film()
    rng_a = ctx.download_rng()
    ctx.upload_rng(scene.rng_map())
    ctx.render_tiles(tile, 0, 3)  # again from scratch: same film, counts not doubled
    b, cb = ctx.film()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(ca, cb) and ca.max() == 3
    assert np.array_equal(rng_a, ctx.download_rng())
    ctx.upload_rng(scene.rng_map())
    ctx.render_tiles(tile, 0, 1)  # 1 + 2 iterations through the same cached graph == 3
    ctx.render_tiles(tile, 1, 2)
    c, cc = ctx.film()
    assert np.array_equal(a.view(np.uint32), c.view(np.uint32)) and np.array_equal(ca, cc)


# ------------------------------------------------------------------ round 2: parity at the benchmarked C5 size
@pytest.fixture(scope="module")
def soup10m():
    scene = prb.Scene.soup(10000000, seed=1234, film=(2048, 2048))
    ctx = make_ctx(scene)
    return scene, ctx, OracleScene(scene)


def test_soup_10m_hits_bit_exact_vs_oracle(soup10m):
    """BASELINE config 5 at its named size (10 M triangles, the scene bench.py --scene c5 times): 64 k primary rays, their
    shadow rays and 64 k incoherent bounce rays through prb_trace_*_device (persistent kernels with the lane-refill path,
    ray streams resident in HBM) against the oracle's own BVH: (entity, prim, u, v, t) and occlusion bit for bit"""
    import torch
    scene, ctx, ora = soup10m
    dev = torch.device("cuda", 0)
    n = 65536
    org, dr, wvl, pix = ctx.generate_camera_rays([(0, 0, 2048, 2048)], 3)
    rs = np.random.RandomState(11)
    sel = np.sort(rs.choice(len(org), n, replace=False))
    org, dr = np.ascontiguousarray(org[sel]), np.ascontiguousarray(dr[sel])

    def device_closest(o, d, tmin=None):
        cols = [torch.from_numpy(np.ascontiguousarray(o[:, i])).to(dev) for i in range(3)] + [torch.from_numpy(np.ascontiguousarray(d[:, i])).to(dev) for i in range(3)]
        tm = torch.from_numpy(tmin).to(dev) if tmin is not None else None
        ent = torch.empty(len(o), dtype=torch.int32, device=dev); prim = torch.empty_like(ent)
        u = torch.empty(len(o), dtype=torch.float32, device=dev); v = torch.empty_like(u); t = torch.empty_like(u)
        ctx.trace_closest_device([c.data_ptr() for c in cols] + [tm.data_ptr() if tm is not None else None, None], len(o),
                                 [ent.data_ptr(), prim.data_ptr(), u.data_ptr(), v.data_ptr(), t.data_ptr()])
        torch.cuda.synchronize(dev)
        return (ent.cpu().numpy().view(np.uint32), prim.cpu().numpy().view(np.uint32), u.cpu().numpy(), v.cpu().numpy(), t.cpu().numpy())

    def device_any(o, d, tmin, tmax):
        cols = [torch.from_numpy(np.ascontiguousarray(o[:, i])).to(dev) for i in range(3)] + [torch.from_numpy(np.ascontiguousarray(d[:, i])).to(dev) for i in range(3)]
        tm, tx = torch.from_numpy(tmin).to(dev), torch.from_numpy(tmax).to(dev)
        occ = torch.empty(len(o), dtype=torch.uint8, device=dev)
        ctx.trace_any_device([c.data_ptr() for c in cols] + [tm.data_ptr(), tx.data_ptr()], len(o), occ.data_ptr())
        torch.cuda.synchronize(dev)
        return occ.cpu().numpy()

    got = device_closest(org, dr)
    ref = ora.trace_closest(org, dr)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), "primary hit ids"
    hit = got[0] != prb.INVALID_ID
    assert hit.mean() > 0.5
    for a, b in zip(got[2:], ref[2:]):
        assert np.array_equal(a[hit].view(np.uint32), b[hit].view(np.uint32))
    # shadow rays towards the point light of the benchmark (Scene::traceShadowRay semantics)
    P = (org[hit] + dr[hit] * got[4][hit, None]).astype(np.float32)
    L = np.array([0, 3, 0], np.float32) - P
    dist = np.linalg.norm(L, axis=1).astype(np.float32)
    L = (L / dist[:, None]).astype(np.float32)
    tmin = np.full(len(P), 1e-4, np.float32)
    tmax = (dist - 1e-3).astype(np.float32)
    occ = device_any(P, L, tmin, tmax)
    assert np.array_equal(occ, ora.trace_any(P, L, tmin, tmax)), "shadow occlusion"
    assert 0.05 < occ.mean() < 1.0
    # incoherent bounce rays
    b = rs.normal(size=P.shape)
    b = (b / np.linalg.norm(b, axis=1, keepdims=True)).astype(np.float32)
    g2 = device_closest(P, b, tmin)
    o2 = ora.trace_closest(P, b, tmin)
    assert np.array_equal(g2[0], o2[0]) and np.array_equal(g2[1], o2[1]), "incoherent hit ids"
    h2 = g2[0] != prb.INVALID_ID
    for a, c in zip(g2[2:], o2[2:]):
        assert np.array_equal(a[h2].view(np.uint32), c[h2].view(np.uint32))


def test_bvh_deeper_than_the_traversal_stack_is_rejected():
    """a BVH whose traversal could overflow the 48-entry stack must not be accepted silently"""
    scene = load_scene("c2_cornellbox")
    d = scene.desc.contents
    n0 = int(d.n_bvh_nodes)
    nodes = np.zeros((n0 + 40, 80), np.uint8)
    nodes[:n0] = np.ctypeslib.as_array(C.cast(d.bvh_nodes, C.POINTER(C.c_uint8)), shape=(n0, 80))
    for k in range(40):  # a chain: every node has one internal child in slot 0 (meta 0x80), the last one is empty
        node = nodes[n0 + k]
        node[:] = 0
        node[24:32] = 0xFF
        if k < 39:
            node[16:20] = np.frombuffer(np.uint32(n0 + k + 1).tobytes(), np.uint8)
            node[24] = 0x80
            node[15] = 1
    saved = (d.bvh_nodes, d.n_bvh_nodes, d.tlas_root)
    try:
        d.bvh_nodes, d.n_bvh_nodes, d.tlas_root = nodes.ctypes.data, n0 + 40, n0
        ctx = prb.Context(0)
        with pytest.raises(prb.PrbError, match="BVH too deep"):
            ctx.upload_scene(scene)
    finally:
        d.bvh_nodes, d.n_bvh_nodes, d.tlas_root = saved


# ------------------------------------------------------------------ round 2: the film combine of the C ABI
def test_film_reduce_single_process_tiles_bit_identical():
    """prb_film_reduce(ctxs, n, PRB_PARTITION_TILES): three contexts own interleaved tiles, the root's film, sample counts,
    AOV sums and feedback bits after the reduce are bit-identical to one context rendering every tile (here the contexts
    share one device; with several GPUs the same kernel reads the peers over NVLink)"""
    from pearray_b200 import multigpu
    src = SKYSUN_ZOO.replace(":camera 'Camera'", ":camera 'Camera' :spectral_hero false")  # leaves feedback bits behind
    scene = prb.Scene.from_string(src)
    scene.settings.want_aov_ext = 1
    tiles = scene.tiles(4, 4)
    spp = 3
    single = make_ctx(scene)
    single.render_tiles(tiles, 0, spp)
    parts = []
    for rank in range(3):
        c = make_ctx(scene)
        c.render_tiles(multigpu.partition_tiles(tiles, rank, 3), 0, spp)
        parts.append(c)
    prb.Context.film_reduce(parts, "tiles")
    a, ca = parts[0].film()
    b, cb = single.film()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(ca, cb)
    assert np.array_equal(parts[0].film_aov().view(np.uint32), single.film_aov().view(np.uint32))
    assert np.array_equal(parts[0].film_aov_ext().view(np.uint32), single.film_aov_ext().view(np.uint32))
    fb = single.film_feedback()
    assert fb.any() and np.array_equal(parts[0].film_feedback(), fb)
    assert parts[0].last_reduce_ms() > 0


def test_film_reduce_carries_the_lpe_channels():
    """the expression channels are films of their own: tiles -> the owner's value, sample ranges -> weighted like the main film"""
    from pearray_b200 import multigpu
    scene = prb.Scene.from_string(LPE_ZOO)
    tiles = scene.tiles(4, 4)
    single = make_ctx(scene)
    single.render_tiles(tiles, 0, 4)
    parts = []
    for rank in range(2):
        c = make_ctx(scene)
        c.render_tiles(multigpu.partition_tiles(tiles, rank, 2), 0, 4)
        parts.append(c)
    prb.Context.film_reduce(parts, "tiles")
    for k in range(len(LPE_EXPRESSIONS)):
        assert np.array_equal(parts[0].film_lpe(k).view(np.uint32), single.film_lpe(k).view(np.uint32)), LPE_EXPRESSIONS[k]
    # sample ranges: 'C.*L' keeps tracking the main film through the weighted combine
    a, b = make_ctx(scene), make_ctx(scene)
    a.render_tiles(tiles, 0, 3)
    b.upload_rng(scene.rng_map(1))
    b.render_tiles(tiles, 3, 5)
    prb.Context.film_reduce([a, b], "samples")
    k = LPE_EXPRESSIONS.index("C.*L")
    np.testing.assert_allclose(a.film_lpe(k), a.film()[0], rtol=1e-5, atol=1e-6)


def test_film_reduce_single_process_sample_ranges():
    """PRB_PARTITION_SAMPLES: two contexts render the iteration ranges [0, 3) and [3, 8) of ONE 8-iteration sequence from
    decorrelated RNG maps; the combined film is the mean over all 8 iterations (weights end_r / total)"""
    from pearray_b200 import multigpu
    scene = load_scene("c0_evaluation")  # block filter radius 0: the download is the unfiltered film
    assert int(scene.settings.filter_radius) == 0
    tile = [(64, 64, 192, 192)]
    a = make_ctx(scene)
    a.render_tiles(tile, 0, 3)
    fa, ca = a.film()
    scene_b = load_scene("c0_evaluation")
    scene_b.settings.seed = multigpu.rank_seed(scene_b.settings.seed, 1)
    b = make_ctx(scene_b)
    b.render_tiles(tile, 3, 5)
    fb, cb = b.film()
    prb.Context.film_reduce([a, b], "samples")
    f, cnt = a.film()
    expect = fa.astype(np.float64) * (3 / 8) + fb.astype(np.float64) * (8 / 8)  # film_r = sum_r / end_r
    assert np.allclose(f, expect, rtol=1e-6, atol=1e-9)
    assert np.array_equal(cnt, ca + cb)
    assert not np.array_equal(fa, fb)


def test_film_reduce_argument_checks():
    scene = load_scene("c2_cornellbox")
    a = make_ctx(scene)
    with pytest.raises(prb.PrbError):
        prb.Context.film_reduce([a, a], "tiles")  # the same context twice
    with pytest.raises(prb.PrbError):
        a.film_reduce_comm("tiles", 4)  # no communicator
    other = make_ctx(load_scene("c3_cornellbox_glassy"))
    with pytest.raises(prb.PrbError):
        prb.Context.film_reduce([a, other], "tiles")  # different film sizes
