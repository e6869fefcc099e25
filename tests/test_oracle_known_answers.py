"""The oracle pinned against the reference's own known-answer tests.

Each test ports one case of /root/reference/src/tests/*.cpp (file:line cited) and runs it against oracle/liboracle.so
(the CPU restatement) -- and, where the product's host library implements the same function, against
libprb200_host.so as well.  Tolerance: PRT_EPSILON = 2 * float epsilon (src/tests/Test.h:224-225) unless the
reference test states its own."""
import ctypes as C

import numpy as np
import pytest

import pearray_b200 as prb
import oracle_binding as ob
from scene_strings import MATERIAL_ZOO, MATERIAL_ZOO2

EPS = 2 * np.finfo(np.float32).eps


def v3(*a):
    v = np.array(a, dtype=np.float32)
    return v


def nrm(*a):
    v = np.array(a, dtype=np.float64)
    return (v / np.linalg.norm(v)).astype(np.float32)


def p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- src/tests/fresnel.cpp:12-60
def test_fresnel_dielectric_one_dot():
    assert abs(ob.lib().orc_fresnel_dielectric(1, 1, 1) - 0) <= EPS


@pytest.mark.parametrize("args,expect", [((0, 1, 1, 1), 1.0), ((1, 1, 1, 1), 0.2), ((0.45, 1, 1, 0), 0.0)])
def test_fresnel_conductor(args, expect):
    assert abs(ob.lib().orc_fresnel_conductor(*args) - expect) <= EPS


@pytest.mark.parametrize("args,expect", [((0, 1, 1), 1.0), ((1, 1, 1), 0.0)])
def test_fresnel_schlick(args, expect):
    assert abs(ob.lib().orc_fresnel_schlick(*args) - expect) <= EPS


# ---------------------------------------------------------------- src/tests/microfacets.cpp:9-93
def test_ggx_iso_pdf_is_d_cos():
    H = nrm(0, 0.2, 0.8)
    D = ob.lib().orc_ndf_ggx_iso(p(H), 0.05)
    assert abs(ob.lib().orc_pdf_ggx_iso(p(H), 0.05) - D * H[2]) <= EPS * max(1.0, abs(D))


def test_ggx_aniso_pdf_is_d_cos():
    H = nrm(0, 0.2, 0.8)
    D = ob.lib().orc_ndf_ggx_aniso(p(H), 0.05, 0.45)
    assert abs(ob.lib().orc_pdf_ggx_aniso(p(H), 0.05, 0.45) - D * H[2]) <= EPS * max(1.0, abs(D))


def test_ggx_iso_equals_aniso():
    H = nrm(0, 0.2, 0.8)
    d1 = ob.lib().orc_ndf_ggx_iso(p(H), 0.05)
    d2 = ob.lib().orc_ndf_ggx_aniso(p(H), 0.05, 0.05)
    # 2 ulp at 0.2126: the port rounds every operation (-ffp-contract=off) whereas the reference is built with
    # -march=native contraction; the two formulas (alpha^2 vs alpha_x*alpha_y) agree to 4 float epsilons
    assert abs(d1 - d2) <= 4 * EPS * max(1.0, abs(d1))


@pytest.mark.parametrize("m", [0.0, 0.245])
def test_microfacet_reflection_reciprocal(m):
    A = nrm(0, 1, 1)
    B = np.zeros(3, np.float32)
    ob.lib().orc_reflect(p(A), p(v3(0, 0, 1)), p(B))
    l = ob.lib()
    a = l.orc_microfacet_reflection_eval(p(A), p(B), m, m, 0, 0)
    b = l.orc_microfacet_reflection_eval(p(B), p(A), m, m, 0, 0)
    assert abs(a - b) <= EPS * max(1.0, abs(a))
    a3 = l.orc_microfacet_reflection_eval_conductor(p(A), p(B), m, m, 0, 0, 0.051585, 3.9046)
    b3 = l.orc_microfacet_reflection_eval_conductor(p(B), p(A), m, m, 0, 0, 0.051585, 3.9046)
    assert abs(a3 - b3) <= EPS * max(1.0, abs(a3))
    a4 = l.orc_microfacet_reflection_pdf(p(A), p(B), m, m, 0, 0)
    b4 = l.orc_microfacet_reflection_pdf(p(B), p(A), m, m, 0, 0)
    assert abs(a4 - b4) <= EPS * max(1.0, abs(a4))


# ---------------------------------------------------------------- src/tests/scattering.cpp:7-47
def test_scattering_reflect_z_equals_general():
    V = nrm(1, 1, 1)
    L2 = np.zeros(3, np.float32)
    ob.lib().orc_reflect(p(V), p(v3(0, 0, 1)), p(L2))
    assert np.allclose(L2, [-V[0], -V[1], V[2]], atol=EPS)  # Scattering::reflect(V) = (-x,-y,z)


def test_scattering_refract():
    V = nrm(1, 1, 1)
    L = np.zeros(3, np.float32)
    tot = C.c_int()
    ob.lib().orc_refract(0.85, p(V), p(v3(0, 0, 1)), p(L), C.byref(tot))
    assert tot.value == 0
    # Snell: sin_t = eta * sin_i, refracted direction on the opposite side of N
    sin_i = np.sqrt(1 - float(V[2]) ** 2)
    assert abs(np.sqrt(L[0] ** 2 + L[1] ** 2) - 0.85 * sin_i) < 1e-6
    assert L[2] < 0 and abs(np.linalg.norm(L) - 1) < 1e-6


def test_scattering_halfway_reflection():
    V, L = nrm(1, 1, 1), nrm(-1, 0, 1)
    H = np.zeros(3, np.float32)
    ob.lib().orc_halfway_reflection(p(V), p(L), p(H))
    assert abs(float(H @ V) - float(H @ L)) <= 4 * EPS


def test_scattering_halfway_transmission():
    n1, n2 = 1.0, 1.55
    V, L = nrm(1, 1, 1), nrm(-1, 0, -1)
    H = np.zeros(3, np.float32)
    ob.lib().orc_halfway_refractive(n1, p(V), n2, p(L), p(H))
    L2 = np.zeros(3, np.float32)
    tot = C.c_int()
    ob.lib().orc_refract(n1 / n2, p(V), p(H), p(L2), C.byref(tot))
    assert np.allclose(L2, L, atol=1e-6)


# ---------------------------------------------------------------- src/tests/sampling.cpp:7-35
def test_cos_hemi_unit_length():
    o = np.zeros(3, np.float32)
    ob.lib().orc_cos_hemi(0.5, 0.5, p(o))
    assert abs(float(o @ o) - 1) <= 4 * EPS
    # cos_hemi(u1,u2): cos(theta) = sqrt(u1)  (src/base/math/Sampling.h:38-51; == power cosine with m = 1)
    assert abs(o[2] - np.sqrt(0.5)) <= 4 * EPS


# ---------------------------------------------------------------- src/tests/tangent.cpp:7-40
@pytest.mark.parametrize("N,Nx,Ny", [((0, 0, 1), (1, 0, 0), (0, 1, 0)), ((0, 1, 0), (1, 0, 0), (0, 0, -1))])
def test_tangent_frame(N, Nx, Ny):
    n = v3(*N)
    x, y = np.zeros(3, np.float32), np.zeros(3, np.float32)
    ob.lib().orc_tangent_frame(p(n), p(x), p(y))
    assert np.allclose(x, Nx, atol=EPS) and np.allclose(y, Ny, atol=EPS)
    assert abs(x @ n) <= EPS and abs(y @ n) <= EPS and abs(x @ y) <= EPS


def test_tangent_frame_orthonormal_random():
    rs = np.random.RandomState(1)
    for _ in range(200):
        n = nrm(*rs.normal(size=3))
        x, y = np.zeros(3, np.float32), np.zeros(3, np.float32)
        ob.lib().orc_tangent_frame(p(n), p(x), p(y))
        assert abs(x @ n) < 1e-6 and abs(y @ n) < 1e-6 and abs(x @ y) < 1e-6
        assert abs(x @ x - 1) < 1e-5 and abs(y @ y - 1) < 1e-5


# ---------------------------------------------------------------- src/tests/distribution.cpp:9-80
def _cdf(values):
    v = np.asarray(values, dtype=np.float32)
    c = np.zeros(len(v) + 1, np.float32)
    for i in range(len(v)):  # Distribution1D::generate: running float sum of f(i)/n, then normalised (Distribution1D.inl:13-40)
        c[i + 1] = c[i] + v[i] / np.float32(len(v))
    integral = c[-1]
    c /= integral
    c[-1] = 1
    return c, float(integral) * len(v)


def test_distribution_pmf():
    c, _ = _cdf([1, 1, 1, 1, 1])
    for u, _i in ((0.05, 0), (0.5, 2), (0.95, 4)):
        pdf = C.c_float()
        i = ob.lib().orc_sample_discrete(p(c), len(c), u, C.byref(pdf))
        assert i == _i and abs(pdf.value - 0.2) <= EPS


def test_distribution_pmf2():
    c, integral = _cdf([1, 2, 3, 4, 5])
    assert abs(integral - 15.0) < 1e-5
    for u, _i in ((0.01, 0), (0.3, 2), (0.9, 4)):
        pdf = C.c_float()
        i = ob.lib().orc_sample_discrete(p(c), len(c), u, C.byref(pdf))
        assert i == _i and abs(pdf.value - (_i + 1) / 15.0) <= 2 * EPS


def test_distribution_continuous_uniform_pdf():
    c, _ = _cdf([1, 1, 1, 1, 1])
    for u in (0.25, 0.5, 0.75):
        pdf = C.c_float()
        x = ob.lib().orc_sample_continuous(p(c), len(c), u, C.byref(pdf))
        assert abs(pdf.value - 1) <= 4 * EPS and abs(x - u) <= 4 * EPS


def test_distribution_continuous_consistency():
    c, _ = _cdf([(i / 16.0) ** 2 for i in range(5)])
    pdf = C.c_float()
    x = ob.lib().orc_sample_continuous(p(c), len(c), 0.5, C.byref(pdf))
    # continuousPdf(x) = (cdf[i+1]-cdf[i]) * n for the bin that contains x (Distribution1D.inl:88-100)
    i = min(int(x * 5), 4)
    assert pdf.value == np.float32((c[i + 1] - c[i]) * np.float32(5))


# ---------------------------------------------------------------- src/tests/random.cpp + pcg32_fast definition
def _pcg32_fast(seed, n):
    """pcg32_fast = mcg_xsh_rs_64_32 (src/core/random/pcg_random.hpp:484-490,812-836,1865): state = seed | 3."""
    s = (seed | 3) & (2 ** 64 - 1)
    out = []
    for _ in range(n):
        old = s
        s = (s * 6364136223846793005) & (2 ** 64 - 1)
        rs = old >> 61
        x = old ^ (old >> 22)
        out.append((x >> (22 + rs)) & 0xFFFFFFFF)
    return out


def test_random_matches_pcg_definition_and_bounds():
    n = 4096
    o32 = np.zeros(n, np.uint32)
    of = np.zeros(n, np.float32)
    ob.lib().orc_random_stream(42, n, p(o32), p(of))
    assert list(o32) == _pcg32_fast(42, n)
    assert (of >= 0).all() and (of < 1).all()
    # Random::getFloat: bits((u32 >> 9) | 0x3F800000) - 1  (src/core/Random.h:133-158)
    exp = ((o32 >> 9) | 0x3F800000).view(np.float32) - np.float32(1)
    assert np.array_equal(of, exp)
    # the product's host Random is the same generator
    h32 = np.zeros(n, np.uint32)
    hf = np.zeros(n, np.float32)
    prb.host_lib().prh_random_stream(42, n, p(h32), p(hf))
    assert np.array_equal(h32, o32) and np.array_equal(hf, of)


# ---------------------------------------------------------------- src/tests/upsampler.cpp:13-123 (golden coefficients + reflectances)
UPSAMPLER_GOLDEN = [
    ((0.8, 0.2, 0.3), (0.000110479, -0.112288, 27.692141), (0.193251, 0.693976, 0.950077, 0.985873, 0.549449)),
    ((0.2, 0.8, 0.4), (-0.000149, 0.156879, -40.710041), (0.783971, 0.299929, 0.013470, 0.010529, 0.120486)),
    ((0.1, 0.3, 0.8), (0.000033, -0.044228, 13.931887), (0.337471, 0.165104, 0.965322, 0.153193, 0.882958)),
    ((1.0, 1.0, 1.0), (0.0, 0.0, 5e6), (1, 1, 1, 1, 1)),
    ((0.0, 0.0, 0.0), (0.0, 0.0, -500.0), (0, 0, 0, 0, 0)),
]
UPSAMPLER_WVL = np.array([532, 615, 346, 720, 416], dtype=np.float32)


@pytest.mark.parametrize("rgb,coeffs,refl", UPSAMPLER_GOLDEN)
def test_upsampler_golden(rgb, coeffs, refl):
    h = prb.host_lib()
    c = np.zeros(3, np.float32)
    assert h.prh_upsample_rgb(p(np.array(rgb, np.float32)), p(c)) == 0, h.prh_last_error()
    assert np.allclose(c, coeffs, atol=1e-4)  # EPS of the reference test
    out = np.zeros(5, np.float32)
    h.prh_upsample_eval(p(c), p(UPSAMPLER_WVL), p(out), 5)
    assert np.allclose(out, refl, atol=1e-4)


def test_upsampler_golden_through_oracle_nodes():
    """the same five colours as 'refl' nodes of a scene, evaluated by the ORACLE's node evaluator"""
    mats = "\n".join("(material :name 'm%d' :type 'diffuse' :albedo (refl %g %g %g))" % (i, *g[0]) for i, g in enumerate(UPSAMPLER_GOLDEN))
    src = """(scene :name 't' :render_width 8 :render_height 8 :camera 'Camera'
      (camera :name 'Camera' :type 'standard' :width 1 :height 1 :local_direction [0,0,-1] :local_up [0,1,0] :local_right [1,0,0] :near 0.1 :far 100)
      %s (entity :name 's' :type 'sphere' :radius 1 :material 'm0'))""" % mats
    scene = prb.Scene.from_string(src)
    ora = ob.OracleScene(scene)
    d = scene.desc.contents
    assert d.n_materials == 5
    for i, (_, _, refl) in enumerate(UPSAMPLER_GOLDEN):
        node = d.materials[i].node[0]
        got = [ora.eval_node(node, float(w)) for w in UPSAMPLER_WVL]
        assert np.allclose(got, refl, atol=1e-4), (i, got, refl)


# ---------------------------------------------------------------- src/tests/materials.cpp:13-165 (eval/pdf/sample self-consistency)
@pytest.fixture(scope="module", params=["zoo", "zoo2"])
def zoo(request):
    scene = prb.Scene.from_string(MATERIAL_ZOO if request.param == "zoo" else MATERIAL_ZOO2)
    return scene, ob.OracleScene(scene)


def _query(scene, mat, V, L=None, seed=42):
    q = (prb.MaterialQuery * 1)()
    q[0].V[:] = [float(x) for x in V]
    if L is not None:
        q[0].L[:] = [float(x) for x in L]
    q[0].wavelength_nm[:] = [560.0, 540.0, 400.0, 600.0]  # materials.cpp:18
    q[0].uv[:] = [0.5, 0.5]
    q[0].ray_flags = prb.device_lib() and 0x01
    q[0].material_id = mat
    q[0].rng_state = seed | 3
    return q


@pytest.mark.parametrize("backside", [False, True])
def test_materials_sample_matches_eval(zoo, backside):
    """sample.PDF_S == eval.PDF_S and sample.IntegralWeight * PDF_S == eval.Weight for the sampled direction
    (materials.cpp:91-137); V = -ray.Direction in shading space with N = +z (constructTestIP :13-37)."""
    scene, ora = zoo
    d = scene.desc.contents
    V = nrm(1, 0, 1) if not backside else -nrm(1, 0, 1)
    for mat in range(d.n_materials):
        only_delta = bool(d.materials[mat].flags & 0x20)
        for seed in (42, 43, 1234567, 99):
            s = ora.material_sample(_query(scene, mat, V, seed=seed))[0]
            if only_delta:
                assert s.flags & 0x2, "hasOnlyDeltaDistribution but the sample is not flagged delta"
                continue
            if s.flags & 0x2:
                continue
            if not any(s.pdf_s):
                continue  # rejected sample
            e = ora.material_eval(_query(scene, mat, V, list(s.L)))[0]
            for k in range(4):
                tol = 2e-4 * max(1.0, abs(e.pdf_s[k]))
                assert abs(s.pdf_s[k] - e.pdf_s[k]) <= tol, (mat, d.materials[mat].type, k, s.pdf_s[k], e.pdf_s[k])
            # IntegralWeight = f cos / pdf[0]  (MaterialData.h:40-79)
            for k in range(4):
                tol = 5e-4 * max(1.0, abs(e.weight[k]))
                assert abs(s.weight[k] * s.pdf_s[0] - e.weight[k]) <= tol, (mat, d.materials[mat].type, k, s.weight[k] * s.pdf_s[0], e.weight[k])


# ---------------------------------------------------------------- mirror.cpp / orennayar.cpp
def _zoo2_material(scene, mtype, nth=0):
    d = scene.desc.contents
    ids = [i for i in range(d.n_materials) if d.materials[i].type == mtype]
    return ids[nth]


def test_mirror_reflects_and_tints():
    scene = prb.Scene.from_string(MATERIAL_ZOO2)
    ora = ob.OracleScene(scene)
    V = nrm(0.3, -0.2, 0.9)
    plain, tinted = _zoo2_material(scene, 6, 0), _zoo2_material(scene, 6, 1)
    s = ora.material_sample(_query(scene, plain, V))[0]
    assert s.flags & 0x2 and list(s.pdf_s) == [1.0] * 4 and list(s.weight) == [1.0] * 4
    assert np.allclose(list(s.L), [-V[0], -V[1], V[2]], atol=0)  # Scattering::reflect in shading space
    t = ora.material_sample(_query(scene, tinted, V))[0]
    node = scene.desc.contents.materials[tinted].node[0]
    assert np.allclose(list(t.weight), [ora.eval_node(node, w) for w in (560.0, 540.0, 400.0, 600.0)], atol=1e-6)


def test_orennayar_roughness_zero_is_lambert_and_energy_is_bounded():
    scene = prb.Scene.from_string(MATERIAL_ZOO2)
    ora = ob.OracleScene(scene)
    d = scene.desc.contents
    ids = [i for i in range(d.n_materials) if d.materials[i].type == 7]
    smooth = [i for i in ids if d.materials[i].f[0] == 0.0][0]
    rough = [i for i in ids if d.materials[i].f[0] == 0.5][0]
    V = nrm(0.5, 0.1, 0.8)
    albedo = [ora.eval_node(d.materials[smooth].node[0], w) for w in (560.0, 540.0, 400.0, 600.0)]
    rs = np.random.RandomState(3)
    n, acc = 4000, np.zeros(4)
    for _ in range(n):
        u1, u2 = rs.rand(2)  # uniform hemisphere directions, pdf 1 / (2 pi)
        L = np.array([np.sqrt(1 - u1 * u1) * np.cos(2 * np.pi * u2), np.sqrt(1 - u1 * u1) * np.sin(2 * np.pi * u2), u1])
        e0 = ora.material_eval(_query(scene, smooth, V, L))[0]
        assert np.allclose(list(e0.weight), np.array(albedo) * L[2] / np.pi, rtol=2e-6, atol=1e-7)
        assert np.allclose(list(e0.pdf_s), [L[2] / np.pi] * 4, rtol=2e-6)
        acc += np.array(list(ora.material_eval(_query(scene, rough, V, L))[0].weight)) * 2 * np.pi
    assert np.all(acc / n < 1.0) and np.all(acc / n > 0.05)  # directional albedo of the rough lobe stays below one
    below = ora.material_eval(_query(scene, rough, V, [0.3, 0.1, -0.9]))[0]
    assert list(below.weight) == [0.0] * 4 and list(below.pdf_s) == [0.0] * 4


# ---------------------------------------------------------------- spectralmapper/cie.cpp
@pytest.mark.parametrize("mapper,channels", [("cie", (0, 1, 2)), ("cie_y", (1,))])
def test_cie_mapper_cdf_is_the_static_cdf_of_the_cie_tables(mapper, channels):
    """the CDF handed to the device is StaticCDF(NM_TO_X + NM_TO_Y + NM_TO_Z) resp. StaticCDF(NM_TO_Y) (CIE.cpp:431-433),
    and wavelengths drawn from it follow that density"""
    src = MATERIAL_ZOO2.replace("(sampler :slot 'aa'", "(spectral_mapper :slot 'pixel' :type '%s') (sampler :slot 'aa'" % mapper)
    scene = prb.Scene.from_string(src)
    d = scene.desc.contents
    m = d.pixel_mapper
    assert (m.type, m.cdf_size, m.trunc_cdf_start, m.trunc_cdf_end) == (3, 442, 0.0, 1.0)
    pool = np.ctypeslib.as_array(d.pool, shape=(d.n_pool,))
    cdf = pool[m.cdf_offset:m.cdf_offset + m.cdf_size]
    h = prb.host_lib()
    h.prh_cie_eval.restype = C.c_float
    h.prh_cie_eval.argtypes = [C.c_int, C.c_float]
    wl = np.arange(390, 831, dtype=np.float32)
    dens = sum(np.array([h.prh_cie_eval(c, float(w)) for w in wl]) for c in channels)
    want = np.concatenate([[0.0], np.cumsum(dens)])
    want /= want[-1]
    assert cdf[0] == 0 and cdf[-1] == 1 and np.all(np.diff(cdf) >= 0)
    np.testing.assert_allclose(cdf, want, atol=2e-5)
    ora = ob.OracleScene(scene)
    _, _, wvl, _ = ora.generate_camera_rays([(0, 0, 32, 32)], 0)
    hist, _ = np.histogram(wvl.ravel(), bins=11, range=(390, 830))
    expect = np.diff(np.interp(np.linspace(0, 1, 12), np.linspace(0, 1, 442), want)) * wvl.size
    assert np.all(np.abs(hist - expect) < 5 * np.sqrt(expect + 1) + 4)


# ---------------------------------------------------------------- sampler/HaltonSampler.cpp
def test_halton_and_hammersley_tables_are_radical_inverses():
    def table(kind, extra):
        src = MATERIAL_ZOO2.replace("(sampler :slot 'aa' :type 'mjitt' :sample_count 16)", "(sampler :slot 'aa' :type '%s' :sample_count 8 %s)" % (kind, extra))
        scene = prb.Scene.from_string(src)
        d = scene.desc.contents
        s = d.aa_sampler
        pool = np.ctypeslib.as_array(d.pool, shape=(d.n_pool,))
        one = pool[s.table_offset:s.table_offset + s.max_samples].copy()
        two = pool[s.table_offset + s.max_samples:s.table_offset + 3 * s.max_samples].reshape(-1, 2).copy()
        return s, one, two

    def radical_inverse(i, b):
        r, f = 0.0, 1.0
        while i > 0:
            f /= b
            r += f * (i % b)
            i //= b
        return r
    s, one, two = table("halton", ":base_x 2 :base_y 3 :burnin 0")
    assert (s.type, s.max_samples, s.m2d_x, s.m2d_y, s.seed) == (5, 8, 2, 3, 0)
    assert np.allclose(one, [0, .5, .25, .75, .125, .625, .375, .875], atol=1e-7)
    assert np.array_equal(two[:, 0], one)
    assert np.allclose(two[:, 1], [radical_inverse(i, 3) for i in range(8)], atol=1e-6)
    s, one, two = table("halton", "")  # defaults: bases 13 / 47, burn-in max(13, 47)
    assert (s.m2d_x, s.m2d_y, s.seed) == (13, 47, 47)
    assert np.allclose(two[:, 0], [radical_inverse(i + 47, 13) for i in range(8)], atol=1e-6)
    s, one, two = table("hammersley", ":base_x 2")
    assert (s.m2d_x, s.m2d_y, s.seed) == (2, 47, 2)  # burn-in defaults to base_x; base 47 only past the table
    assert np.allclose(two[:, 1], (0.5 + np.arange(8)) / 8, atol=1e-7)
    assert np.allclose(two[:, 0], [radical_inverse(i + 2, 2) for i in range(8)], atol=1e-7)


# ---------------------------------------------------------------- materials/blend.cpp, add.cpp
def test_blend_and_add_follow_their_children():
    from scene_strings import MATERIAL_ZOO3
    scene = prb.Scene.from_string(MATERIAL_ZOO3)
    ora = ob.OracleScene(scene)
    d = scene.desc.contents
    mats = {}
    for i in range(d.n_materials):
        mats.setdefault((d.materials[i].type, d.materials[i].flags & 0x300), []).append(i)
    V, L = nrm(0.3, 0.2, 0.9), nrm(-0.4, 0.1, 0.8)
    ev = lambda m: ora.material_eval(_query(scene, m, V, L))[0]
    arr = lambda x: np.array(list(x), np.float32)
    # blend, no delta child: (1 - f) * m0 + f * m1 for weight and pdf (blend.cpp:66-76)
    b = mats[(8, 0)][0]
    m0, m1, f = d.materials[b].node[0], d.materials[b].node[1], np.float32(d.materials[b].f[0])
    e, e0, e1 = ev(b), ev(m0), ev(m1)
    assert np.array_equal(arr(e.weight), (1 - f) * arr(e0.weight) + f * arr(e1.weight))
    assert np.array_equal(arr(e.pdf_s), (1 - f) * arr(e0.pdf_s) + f * arr(e1.pdf_s))
    # blend, first child delta: the second child scaled by f (blend.cpp:53-58); second child delta: the first by 1 - f
    b = mats[(8, 0x100)][0]
    f = np.float32(d.materials[b].f[0])
    e, e1 = ev(b), ev(d.materials[b].node[1])
    assert np.array_equal(arr(e.weight), arr(e1.weight) * f) and np.array_equal(arr(e.pdf_s), arr(e1.pdf_s) * f)
    b = mats[(8, 0x200)][0]
    f = np.float32(d.materials[b].f[0])
    e, e0 = ev(b), ev(d.materials[b].node[0])
    assert np.array_equal(arr(e.weight), arr(e0.weight) * (1 - f)) and np.array_equal(arr(e.pdf_s), arr(e0.pdf_s) * (1 - f))
    # both delta: only-delta material, its samples are delta
    b = mats[(8, 0x300)][0]
    assert d.materials[b].flags & 0x20
    assert ora.material_sample(_query(scene, b, V))[0].flags & 0x2
    # add: weights add, pdfs average (add.cpp:53-62); with a delta child the other one with half the pdf (:47-52)
    a = mats[(9, 0)][0]
    e, e0, e1 = ev(a), ev(d.materials[a].node[0]), ev(d.materials[a].node[1])
    assert np.array_equal(arr(e.weight), arr(e0.weight) + arr(e1.weight))
    assert np.array_equal(arr(e.pdf_s), (arr(e0.pdf_s) + arr(e1.pdf_s)) / 2)
    a = mats[(9, 0x100)][0]
    e, e1 = ev(a), ev(d.materials[a].node[1])
    assert np.array_equal(arr(e.weight), arr(e1.weight)) and np.array_equal(arr(e.pdf_s), arr(e1.pdf_s) * np.float32(0.5))
    # sampling draws one number to pick a child, samples it with the following numbers and scales by its share (blend.cpp:106-122)
    h = prb.host_lib()
    h.prh_random_advance.restype = C.c_uint64
    h.prh_random_advance.argtypes = [C.c_uint64, C.c_uint64]
    b = mats[(8, 0)][0]
    f = np.float32(d.materials[b].f[0])
    picked = []
    for seed in [(0x9E3779B97F4A7C15 * k) & 0xFFFFFFFFFFFFFFFF for k in range(1, 41)]:  # well mixed states (an MCG outputs ~0 for tiny ones)
        s = ora.material_sample(_query(scene, b, V, seed=seed))[0]
        matches = []
        for k, share in ((0, 1 - f), (1, f)):
            q = _query(scene, d.materials[b].node[k], V, seed=seed)
            q[0].rng_state = h.prh_random_advance(seed | 3, 1)
            c = ora.material_sample(q)[0]
            if np.array_equal(arr(s.L), arr(c.L)) and np.array_equal(arr(s.weight), arr(c.weight) * share) and np.array_equal(arr(s.pdf_s), arr(c.pdf_s) * share):
                matches.append(k)
        assert len(matches) == 1, (seed, matches)
        picked.append(matches[0])
    assert 0 < sum(picked) < len(picked)  # both children get picked; f = 0.3 -> mostly the first
    assert sum(picked) < len(picked) / 2


# ---------------------------------------------------------------- spectralmapper/agh.cpp
@pytest.mark.parametrize("cmis", [True, False])
def test_agh_mapper_follows_its_closed_form(cmis):
    """lambda = B - atanh(C - N u) / A with A = 0.0072, B = 538 (agh.cpp:19-36): wavelengths stay inside the camera range, are
    densest around B, and the pdf handed on is 1 / (cosh^2(A (lambda - B)) N); the hero form rotates one sample by span / 4"""
    extra = "(spectral_mapper :slot 'pixel' :type 'agh'%s)" % ("" if cmis else " :cmis false")
    scene = prb.Scene.from_string(MATERIAL_ZOO2.replace("(sampler :slot 'aa'", extra + " (sampler :slot 'aa'"))
    m = scene.desc.contents.pixel_mapper
    A, B = 0.0072, 538.0
    C, N = np.tanh(A * (B - 390.0)), np.tanh(A * (B - 390.0)) - np.tanh(A * (B - 830.0))
    assert m.type == (4 if cmis else 5)
    assert abs(m.trunc_cdf_start - C) < 1e-6 and abs(m.trunc_cdf_end - N) < 1e-6
    ora = ob.OracleScene(scene)
    _, _, wvl, _ = ora.generate_camera_rays([(0, 0, 32, 32)], 0)
    assert wvl.min() >= 390.0 and wvl.max() <= 830.0
    hero = wvl[:, 0] if not cmis else wvl.ravel()
    inner = np.mean(np.abs(hero - B) < 70)  # P(|lambda - B| < 70) = 2 tanh(0.504) / N = 0.529
    assert abs(inner - 2 * np.tanh(A * 70) / N) < 0.06
    if not cmis:
        span = 440.0
        for k in range(1, 4):
            assert np.allclose(wvl[:, k], 390.0 + np.mod(wvl[:, 0] - 390.0 + k * span / 4, span), atol=2e-3)


def test_nested_blend_and_add():
    """children of a blend / add may be blends / adds themselves: the nested result is the composition of the one-level rules"""
    from scene_strings import MATERIAL_ZOO4
    scene = prb.Scene.from_string(MATERIAL_ZOO4)
    ora = ob.OracleScene(scene)
    d = scene.desc.contents
    V, L = nrm(0.3, 0.2, 0.9), nrm(-0.4, 0.1, 0.8)
    ev = lambda m: ora.material_eval(_query(scene, m, V, L))[0]
    arr = lambda x: np.array(list(x), np.float32)
    combos = [i for i in range(d.n_materials) if d.materials[i].type in (8, 9)]
    nested = [i for i in combos if d.materials[d.materials[i].node[0]].type in (8, 9) or d.materials[d.materials[i].node[1]].type in (8, 9)]
    assert len(nested) == 4
    n2, n3, n_delta_first, n_all = nested
    half = np.float32(0.5)
    # n2 = blend(b_none, m_oren, 0.5): (1 - f) * blend(...) + f * oren
    m = d.materials[n2]
    e, e0, e1 = ev(n2), ev(m.node[0]), ev(m.node[1])
    assert d.materials[m.node[0]].type == 8
    assert np.array_equal(arr(e.weight), (1 - half) * arr(e0.weight) + half * arr(e1.weight))
    assert np.array_equal(arr(e.pdf_s), (1 - half) * arr(e0.pdf_s) + half * arr(e1.pdf_s))
    # n3 = add(n2, a_first): three levels; weights add, pdfs average
    m = d.materials[n3]
    e, e0, e1 = ev(n3), ev(m.node[0]), ev(m.node[1])
    assert np.array_equal(arr(e.weight), arr(e0.weight) + arr(e1.weight)) and np.array_equal(arr(e.pdf_s), (arr(e0.pdf_s) + arr(e1.pdf_s)) / 2)
    # n_delta_first = blend(b_all [all delta], n2, 0.7): MaterialDelta::First -> the second (nested) child scaled by f
    m = d.materials[n_delta_first]
    assert m.flags & 0x100 and not m.flags & 0x200
    f = np.float32(m.f[0])
    e, e1 = ev(n_delta_first), ev(m.node[1])
    assert np.array_equal(arr(e.weight), arr(e1.weight) * f) and np.array_equal(arr(e.pdf_s), arr(e1.pdf_s) * f)
    # n_all: every leaf below is delta -> only-delta material, its samples are delta
    assert d.materials[n_all].flags & 0x20
    assert ora.material_sample(_query(scene, n_all, V))[0].flags & 0x2
    # sampling n3 draws one number per level; whatever leaf is reached, the pdf carries the product of the shares: compare with the
    # pdf the eval of the same direction reports for non-delta samples of a tree without delta leaves (n2)
    for seed in [(0x9E3779B97F4A7C15 * k) & 0xFFFFFFFFFFFFFFFF for k in range(1, 21)]:
        s = ora.material_sample(_query(scene, n2, V, seed=seed))[0]
        if s.flags & 0x2 or not np.all(np.isfinite(arr(s.pdf_s))) or arr(s.pdf_s)[0] <= 0:
            continue
        shares = {np.float32(0.5) * np.float32(0.7), np.float32(0.5) * np.float32(0.3), np.float32(0.5)}  # b_none: 0.7 / 0.3, n2: 0.5
        assert any(np.isclose(arr(s.pdf_s)[0] / sh, arr(ora.material_eval(_query(scene, leaf, V, arr(s.L)))[0].pdf_s)[0], rtol=1e-5)
                   for sh in shares for leaf in range(d.n_materials) if d.materials[leaf].type not in (8, 9)), seed
    # four levels are refused by the loader (the device unrolls three)
    deep = MATERIAL_ZOO4.replace(" (entity :name 'floor'", " (material :name 'n4' :type 'blend' :material1 'n3' :material2 'm_diffuse')\n (entity :name 'floor'")
    deep = deep.replace(":material 'a_none' :position", ":material 'n4' :position")
    sc4 = prb.Scene.from_string(deep)
    d4 = sc4.desc.contents
    assert d4.n_materials == d.n_materials  # n4 was not created (logged as an error)


def test_depth_of_field_rays_focus_on_the_focal_plane():
    """PerspectiveCamera<HasDOF> (perspective.cpp:45-113): with the same seeds a lens camera draws the same film and lens samples
    as the pinhole camera; its ray starts inside the aperture disc and crosses the pinhole ray of the same sample on the focal
    plane, (fstop + 1) x |direction| in front of the camera"""
    from scene_strings import DOF_ZOO
    fstop, apr = 3.0, 0.08
    pin = prb.Scene.from_string(DOF_ZOO.format(dof=""))
    dof = prb.Scene.from_string(DOF_ZOO.format(dof=":fstop %g :aperture_radius %g" % (fstop, apr)))
    assert pin.desc.contents.camera.has_dof == 0 and dof.desc.contents.camera.has_dof == 1
    tiles = [(0, 0, pin.width, pin.height)]
    po, pd, pw, _ = ob.OracleScene(pin).generate_camera_rays(tiles, 2)
    do, dd, dw, _ = ob.OracleScene(dof).generate_camera_rays(tiles, 2)
    assert np.array_equal(pw, dw)  # same wavelengths: the random streams stay in step
    cam = np.array([0, 0, 4.0])
    fwd = np.array([0, 0, -1.0])
    assert np.allclose(po, cam)
    off = do.astype(np.float64) - cam
    assert np.abs(off @ fwd).max() < 1e-6 and np.linalg.norm(off, axis=1).max() <= apr * 1.0001 and np.linalg.norm(off, axis=1).max() > 0.5 * apr
    depth = fstop + 1  # |local_direction| = 1
    on_plane_pin = cam + pd.astype(np.float64) * (depth / (pd.astype(np.float64) @ fwd))[:, None]
    t = (depth - off @ fwd) / (dd.astype(np.float64) @ fwd)
    on_plane_dof = do.astype(np.float64) + dd.astype(np.float64) * t[:, None]
    assert np.abs(on_plane_dof - on_plane_pin).max() < 2e-5


# ---------------------------------------------------------------- src/tests/sphere.cpp:11-68 (Spherical), :77-110 (Sphere); plane.cpp:66-122
GEOMETRY_SCENE = """
(scene :name 'geometry' :render_width 8 :render_height 8 :camera 'Camera'
 (integrator :type 'direct' :max_ray_depth 2)
 (sampler :slot 'aa' :type 'random' :sample_count 1)
 (camera :name 'Camera' :type 'standard' :width 1 :height 1 :local_direction [0,0,-1] :local_up [0,1,0] :local_right [1,0,0]
   :near 0.01 :far 100 :transform [1,0,0,0, 0,1,0,0, 0,0,1,30, 0,0,0,1])
 (material :name 'white' :type 'diffuse' :albedo 1)
 %s
 (light :type 'env' :radiance 1)
)
"""
# the reference computes these with libm's sin / cos / atan2 / acos; both sides here use the correctly rounded functions (DESIGN.md
# section 3), a few ulp apart: 4 x PRT_EPSILON
SPHERICAL_TOL = 4 * EPS


@pytest.mark.parametrize("uv,expect", [((0.0, 0.0), (0.0, 0.0)), ((1.0, 0.0), (0.0, 0.0)), ((0.0, 1.0), (0.5, 1.0)), ((1.0, 1.0), (0.5, 1.0)),
                                       ((0.5, 0.5), (0.5, 0.5)), ((0.75, 0.25), (0.75, 0.25)), ((0.25, 0.75), (0.25, 0.75))])
def test_spherical_uv_round_trip(uv, expect):
    """Spherical::uv_from_normal(cartesian_from_uv(u, v)), src/tests/sphere.cpp:12-67 incl. the ambiguous poles / seam"""
    n = np.zeros(3, np.float32)
    got = np.zeros(2, np.float32)
    ob.lib().orc_cartesian_from_uv(uv[0], uv[1], p(n))
    assert abs(float(np.linalg.norm(n)) - 1) < SPHERICAL_TOL
    ob.lib().orc_uv_from_normal(p(n), p(got))
    # the seam u = 0 == u = 1: compare modulo 1 there
    du = abs(float(got[0]) - expect[0])
    assert min(du, abs(du - 1)) < SPHERICAL_TOL and abs(float(got[1]) - expect[1]) < SPHERICAL_TOL


def test_sphere_intersections():
    """Sphere::intersects, src/tests/sphere.cpp:77-110: from outside (t = 1), pointing away (miss), from inside (t = 1)"""
    scene = prb.Scene.from_string(GEOMETRY_SCENE % "(entity :name 'ball' :type 'sphere' :radius 1 :material 'white')")
    ora = ob.OracleScene(scene)
    org = np.array([[-2, 0, 0], [-2, 0, 0], [0, 0, 0]], np.float32)
    dr = np.array([[1, 0, 0], [-1, 0, 0], [1, 0, 0]], np.float32)
    ent, prim, u, v, t = ora.trace_closest(org, dr, tmin=np.zeros(3, np.float32))
    assert ent[0] == 0 and abs(float(t[0]) - 1) <= EPS
    assert ent[1] == 0xFFFFFFFF
    assert ent[2] == 0 and abs(float(t[2]) - 1) <= EPS


@pytest.mark.parametrize("axes,origin,direction,expect", [
    (("[1,0,0]", "[0,1,0]"), (0.5, 0.5, -1), (0, 0, 1), (1.0, 0.5, 0.5)),      # plane.cpp:66-78  "Intersects 1"
    (("[1,0,0]", "[0,1,0]"), (0.5, 0.5, -1), (0, 1, 0), None),                   # plane.cpp:80-89  "Intersects 2": parallel, no hit
    (("[10,0,0]", "[0,10,0]"), (5, 5, -1), (0, 0, 1), (1.0, 0.5, 0.5)),         # plane.cpp:91-103 "Intersects 3"
    (("[10,0,0]", "[0,20,0]"), (5, 10, -1), (0, 0, 1), (1.0, 0.5, 0.5)),        # plane.cpp:105-117 "Intersects 4"
])
def test_plane_intersections(axes, origin, direction, expect):
    """Plane::intersects: hit distance and the plane parameters (u, v) of the hit, src/tests/plane.cpp:66-117 -- here through the
    two triangles the plane entity is traced as (Embree's quad convention)"""
    scene = prb.Scene.from_string(GEOMETRY_SCENE % ("(entity :name 'quad' :type 'plane' :x_axis %s :y_axis %s :material 'white')" % axes))
    ora = ob.OracleScene(scene)
    ent, prim, u, v, t = ora.trace_closest(np.array([origin], np.float32), np.array([direction], np.float32), tmin=np.zeros(1, np.float32))
    if expect is None:
        assert ent[0] == 0xFFFFFFFF
    else:
        assert ent[0] == 0
        assert abs(float(t[0]) - expect[0]) <= EPS and abs(float(u[0]) - expect[1]) <= EPS and abs(float(v[0]) - expect[2]) <= EPS
