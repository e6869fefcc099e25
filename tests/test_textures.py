"""Image textures and the image-based environment light: NonParametricImageNode
(reference src/loader/shader/ImageNode.cpp:93-182) and EnvironmentLight<UseDistribution = true>
(plugins/main/infinitelights/environment.cpp:54-104,172-198).  The texel fetch is OpenImageIO's in the reference (not in this
image): what is pinned here are known answers that hold for any correct implementation -- a lookup at a texel centre returns
that texel, a bilinear lookup half way between two centres their mean, t = 1 - v, the wrap modes, sRGB files are linearised with
RGBConverter::linearize as written -- each pushed through the reference's own RGB -> spectrum upsampling
(SpectralUpsampler::prepare / compute, pinned to the reference's golden coefficients in test_oracle_known_answers.py)."""
import ctypes as C
import struct
import zlib

import numpy as np
import pytest

import pearray_b200 as prb
from oracle_binding import OracleScene

WVL = np.array([450.0, 520.0, 600.0, 680.0], np.float32)


def write_pfm(path, img):
    h, w, c = img.shape
    with open(path, "wb") as f:
        f.write(b"PF\n%d %d\n-1.0\n" % (w, h))
        f.write(np.ascontiguousarray(img[::-1], "<f4").tobytes())  # rows bottom to top


def write_ppm(path, img8):
    h, w, _ = img8.shape
    with open(path, "wb") as f:
        f.write(b"P6\n# a comment\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(img8, np.uint8).tobytes())


def write_exr_zip(path, img, half=False):
    """single-part scanline OpenEXR, ZIP compression (16-line blocks: predictor + byte interleave + zlib), channels B G R"""
    h, w, _ = img.shape
    names = ["B", "G", "R"]
    ch = b"".join(n.encode() + b"\0" + struct.pack("<i", 1 if half else 2) + b"\0\0\0\0" + struct.pack("<ii", 1, 1) for n in names) + b"\0"

    def attr(name, typ, val):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(val)) + val

    hdr = struct.pack("<ii", 20000630, 2)
    hdr += attr("channels", "chlist", ch) + attr("compression", "compression", b"\3")
    hdr += attr("dataWindow", "box2i", struct.pack("<4i", 0, 0, w - 1, h - 1)) + attr("displayWindow", "box2i", struct.pack("<4i", 0, 0, w - 1, h - 1))
    hdr += attr("lineOrder", "lineOrder", b"\0") + attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    hdr += attr("screenWindowCenter", "v2f", struct.pack("<ff", 0, 0)) + attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0"
    blocks = []
    for y0 in range(0, h, 16):
        raw = b""
        for y in range(y0, min(h, y0 + 16)):
            for k in (2, 1, 0):  # B, G, R planes of the row
                raw += img[y, :, k].astype("<f2" if half else "<f4").tobytes()
        a = np.frombuffer(raw, np.uint8)
        t = np.concatenate([a[0::2], a[1::2]]).astype(np.int32)  # interleave: even bytes first
        p = t.copy()
        p[1:] = (t[1:] - t[:-1] + 128 + 256) % 256  # predictor
        comp = zlib.compress(p.astype(np.uint8).tobytes())
        data = comp if len(comp) < len(raw) else raw
        blocks.append(struct.pack("<ii", y0, len(data)) + data)
    off = len(hdr) + 8 * len(blocks)
    table = b""
    for b in blocks:
        table += struct.pack("<Q", off)
        off += len(b)
    with open(path, "wb") as f:
        f.write(hdr + table + b"".join(blocks))


def upsampled(rgb):
    h = prb.host_lib()
    c, out = np.zeros(3, np.float32), np.zeros(4, np.float32)
    assert h.prh_upsample_rgb(np.asarray(rgb, np.float32).ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p)) == 0
    h.prh_upsample_eval(c.ctypes.data_as(C.c_void_p), WVL.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), 4)
    return out


SCENE = """
(scene :name 'tex' :render_width 32 :render_height 32 :camera 'Camera'
 (integrator :type 'direct' :max_ray_depth 4)
 (sampler :slot 'aa' :type 'mjitt' :sample_count 16)
 (camera :name 'Camera' :type 'standard' :width 1 :height 1 :local_direction [0,0,-1] :local_up [0,1,0] :local_right [1,0,0]
   :near 0.1 :far 100 :transform [1,0,0,0, 0,1,0,0, 0,0,1,4, 0,0,0,1])
 (texture :name 'albedo' :type 'color' :file '%(file)s' %(options)s)
 (material :name 'm' :type 'diffuse' :albedo (texture 'albedo'))
 (entity :name 'floor' :type 'plane' :centering true :width 3 :height 3 :material 'm' :position [0,0,0])
 (light :type 'env' :radiance %(env)s)
)
"""


def node_values(scene, uv):
    """the albedo node of material 0 at the given surface parameters, four wavelengths each, through the oracle"""
    ora = OracleScene(scene)
    node = scene.desc.contents.materials[0].node[0]
    return np.array([[ora.eval_node(node, float(w), float(u), float(v)) for w in WVL] for u, v in uv], np.float32)


def test_texel_centres_bilinear_midpoints_and_wraps(tmp_path):
    rs = np.random.RandomState(3)
    img = (0.05 + 0.9 * rs.rand(4, 6, 3)).astype(np.float32)  # 6 wide, 4 high, rows top to bottom
    write_pfm(str(tmp_path / "t.pfm"), img)
    H, W = img.shape[:2]
    centres = [((x + 0.5) / W, 1 - (y + 0.5) / H) for y in range(H) for x in range(W)]  # t = 1 - v
    want = np.array([upsampled(img[y, x]) for y in range(H) for x in range(W)])
    for interp in ("closest", "bilinear"):
        scene = prb.Scene.from_string(SCENE % dict(file=tmp_path / "t.pfm", options=":interpolation '%s'" % interp, env="1"))
        assert scene.desc.contents.upsampler_res == 64
        np.testing.assert_allclose(node_values(scene, centres), want, rtol=0, atol=2e-6, err_msg=interp)
    # half way between two texel centres: the mean of the two (bilinear), exactly the left / right one either side of it (closest)
    scene = prb.Scene.from_string(SCENE % dict(file=tmp_path / "t.pfm", options=":interpolation 'bilinear'", env="1"))
    mid = node_values(scene, [(2.0 / W, 1 - 0.5 / H)])[0]
    np.testing.assert_allclose(mid, upsampled(0.5 * (img[0, 1] + img[0, 2])), atol=2e-6)
    # the default interpolation (smart bicubic without derivatives -> B-spline bicubic) reproduces a constant image
    const = np.full((5, 5, 3), 0.3, np.float32) * np.array([1.0, 0.5, 2.0], np.float32)
    write_pfm(str(tmp_path / "c.pfm"), const)
    scene = prb.Scene.from_string(SCENE % dict(file=tmp_path / "c.pfm", options=":wrap 'clamp'", env="1"))
    np.testing.assert_allclose(node_values(scene, [(0.37, 0.61), (0.02, 0.98)]), np.tile(upsampled(const[0, 0]), (2, 1)), atol=1e-5)
    # wrap modes at u = 1.25 (a quarter past the right edge), closest: black -> zero reflectance spectrum, clamp -> last column,
    # periodic -> u = 0.25, mirror -> u = 0.75
    v0 = 1 - 0.5 / H
    col = lambda u: int(u * W)
    cases = {"black": upsampled([0, 0, 0]), "clamp": upsampled(img[0, W - 1]), "periodic": upsampled(img[0, col(0.25)]), "mirror": upsampled(img[0, col(0.75)])}
    for wrap, expect in cases.items():
        scene = prb.Scene.from_string(SCENE % dict(file=tmp_path / "t.pfm", options=":interpolation 'closest' :wrap '%s'" % wrap, env="1"))
        np.testing.assert_allclose(node_values(scene, [(1.25, v0)])[0], expect, atol=2e-6, err_msg=wrap)


def test_image_formats_decode_to_the_same_texels(tmp_path):
    rs = np.random.RandomState(5)
    img = (0.05 + 0.9 * rs.rand(20, 7, 3)).astype(np.float32)  # 20 rows: two ZIP blocks
    write_pfm(str(tmp_path / "a.pfm"), img)
    write_exr_zip(str(tmp_path / "a.exr"), img)
    write_exr_zip(str(tmp_path / "h.exr"), img, half=True)
    H, W = img.shape[:2]
    uv = [((x + 0.5) / W, 1 - (y + 0.5) / H) for y in (0, 7, 19) for x in (0, 3, 6)]
    vals = {}
    for name in ("a.pfm", "a.exr", "h.exr"):
        scene = prb.Scene.from_string(SCENE % dict(file=tmp_path / name, options=":interpolation 'closest'", env="1"))
        vals[name] = node_values(scene, uv)
    assert np.array_equal(vals["a.pfm"], vals["a.exr"])
    half = img.astype(np.float16).astype(np.float32)
    want = np.array([upsampled(half[y, x]) for y in (0, 7, 19) for x in (0, 3, 6)])
    np.testing.assert_allclose(vals["h.exr"], want, atol=2e-6)
    # 8-bit PPM: sRGB encoded -> RGBConverter::linearize AS WRITTEN in the reference (x / 12.92 * x below 0.04045)
    img8 = rs.randint(0, 256, size=(3, 4, 3)).astype(np.uint8)
    img8[0, 0] = (5, 128, 250)
    write_ppm(str(tmp_path / "s.ppm"), img8)
    scene = prb.Scene.from_string(SCENE % dict(file=tmp_path / "s.ppm", options=":interpolation 'closest'", env="1"))
    x = img8[0, 0].astype(np.float32) / np.float32(255)
    lin = np.where(x <= 0.04045, x / np.float32(12.92) * x, ((x + 0.055) / 1.055) ** 2.4).astype(np.float32)
    np.testing.assert_allclose(node_values(scene, [(0.5 / 4, 1 - 0.5 / 3)])[0], upsampled(lin), atol=3e-6)
    # an unreadable file is logged as an error and the texture is not created (TextureParser.cpp:168-171); the material then
    # falls back to its default albedo like in the reference -- never a silently black image node
    scene = prb.Scene.from_string(SCENE % dict(file=tmp_path / "missing.exr", options="", env="1"))
    d = scene.desc.contents
    assert all(d.nodes[i].type != 7 for i in range(d.n_nodes)) and d.upsampler_res == 0


def env_scene(tmp_path, distribution=True):
    """a 16 x 8 latitude-longitude map: dim everywhere, one bright 2 x 2 patch"""
    img = np.full((8, 16, 3), 0.02, np.float32)
    img[2:4, 5:7] = (0.9, 0.8, 0.7)
    write_pfm(str(tmp_path / "env.pfm"), img)
    opts = "" if distribution else ":distribution false"
    src = SCENE.replace("(light :type 'env' :radiance %(env)s)", "(texture :name 'sky' :type 'color' :file '%s' :interpolation 'closest' :wrap 'periodic')\n"
                        " (light :type 'env' :radiance (texture 'sky') %s)" % (tmp_path / "env.pfm", opts))
    return prb.Scene.from_string(src % dict(file=tmp_path / "env.pfm", options="", env="1")), img


def test_environment_map_distribution(tmp_path):
    """EnvironmentLightFactory::create builds a Distribution2D over the image (sin(theta) x max of four preset wavelengths);
    sampleDir draws (u, v) from it, eval returns the same pdf for that direction and the pdf integrates to one over the sphere"""
    scene, img = env_scene(tmp_path)
    d = scene.desc.contents
    assert d.n_lights == 1 and d.lights[0].type == 1 and (d.lights[0].dist_w, d.lights[0].dist_h) == (16, 8)
    pool = np.ctypeslib.as_array(d.pool, (d.n_pool,))
    l = d.lights[0]
    marginal = pool[l.dist_offset:l.dist_offset + 9]
    assert marginal[0] == 0 and abs(marginal[-1] - 1) < 1e-6 and np.all(np.diff(marginal) >= 0)
    rows = np.diff(marginal)
    # the bright image rows 2 and 3 (from the top) are distribution rows 5 and 4: v runs bottom to top (t = 1 - v)
    assert rows[4] + rows[5] > 0.5 and rows[2] + rows[3] < 0.2
    ora = OracleScene(scene)
    state, inv_pdf, bright = 0x853C49E6748FEA9B | 3, [], 0
    for _ in range(4000):
        r, state = ora.light_sample_and_eval(0, (0.0, 0.0, 0.5), [560.0, 540.0, 400.0, 600.0], state)
        assert abs(np.linalg.norm(r["outgoing"]) - 1) < 1e-5 and r["pdf"] > 0
        assert abs(r["pdf"] - r["eval_pdf"]) <= 2e-3 * r["pdf"] and np.allclose(r["radiance"], r["eval_radiance"], rtol=2e-3, atol=1e-6)
        inv_pdf.append(1.0 / r["pdf"])
        bright += r["radiance"][0] > 0.5
    assert abs(np.mean(inv_pdf) - 4 * np.pi) < 0.1 * 4 * np.pi  # E[1 / pdf] = measure of the sphere
    assert bright > 0.5 * len(inv_pdf)  # importance sampling finds the patch (1.6 % of the map)
    # :distribution false -> the cosine-hemisphere branch
    plain, _ = env_scene(tmp_path, distribution=False)
    assert plain.desc.contents.lights[0].dist_w == 0


def test_textured_scene_renders_on_the_oracle(tmp_path):
    scene, _ = env_scene(tmp_path)
    r = OracleScene(scene).render([(0, 0, 32, 32)], 0, 4)
    assert np.isfinite(r["film"]).all() and r["film"].max() > 0 and not r["feedback"].any()
    assert r["stats"]["shadow_ray_count"] > 0
