"""Output specification + EXR writer of the host layer (reference src/loader/output/io/OutputSpecification.cpp,
ImageWriter.cpp): channel naming and ordering, tone mapping, AOV weighting, data/display windows, and that an independent
EXR reader (OpenCV, when it was built with OpenEXR) decodes the file."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

import pearray_b200 as prb

SCENE = """
(scene :name 'out' :render_width 24 :render_height 16 :camera 'Camera'
 (integrator :type 'direct' :max_ray_depth 2)
 (sampler :slot 'aa' :type 'random' :sample_count 4)
 (output :name 'image' (channel :type 'color' :color 'srgb'))
 (output :name 'aovs'
   (channel :type 'rgb' :color 'xyz') (channel :type 'depth') (channel :type 'n') (channel :type 'p') (channel :type 'uv')
   (channel :type 'id') (channel :type 'samples') (channel :type 'ng') (channel :type 'color' :color 'lum' :lpe 'C.*L'))
 (camera :name 'Camera' :type 'standard' :width 1 :height 1 :local_direction [0,0,-1] :local_up [0,1,0] :local_right [1,0,0]
   :near 0.1 :far 100 :transform [1,0,0,0, 0,1,0,0, 0,0,1,4, 0,0,0,1])
 (material :name 'm' :type 'diffuse' :albedo 0.5)
 (entity :name 's' :type 'sphere' :radius 1 :material 'm')
 (light :type 'env' :radiance (illuminant "D65"))
)
"""


def read_exr(path):
    """minimal reader for single-part, uncompressed scanline files: returns (attributes, {channel: HxW float32})"""
    raw = open(path, "rb").read()
    assert struct.unpack_from("<ii", raw, 0) == (20000630, 2)
    pos, attrs = 8, {}

    def cstr(p):
        e = raw.index(b"\0", p)
        return raw[p:e].decode(), e + 1
    while raw[pos] != 0:
        name, pos = cstr(pos)
        typ, pos = cstr(pos)
        size = struct.unpack_from("<i", raw, pos)[0]
        attrs[name] = (typ, raw[pos + 4:pos + 4 + size])
        pos += 4 + size
    pos += 1
    chans, p, body = [], 0, attrs["channels"][1]
    while body[p] != 0:
        e = body.index(b"\0", p)
        nm = body[p:e].decode()
        ptype, = struct.unpack_from("<i", body, e + 1)
        xs, ys = struct.unpack_from("<ii", body, e + 9)
        assert ptype == 2 and (xs, ys) == (1, 1)
        chans.append(nm)
        p = e + 17
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    offsets = struct.unpack_from("<%dQ" % h, raw, pos)
    planes = {c: np.empty((h, w), np.float32) for c in chans}
    for y in range(h):
        yy, size = struct.unpack_from("<ii", raw, offsets[y])
        assert yy == y0 + y and size == len(chans) * w * 4
        row = np.frombuffer(raw, "<f4", len(chans) * w, offsets[y] + 8).reshape(len(chans), w)
        for k, c in enumerate(chans):
            planes[c][y] = row[k]
    return attrs, chans, planes


def _film(w, h, seed=5):
    rs = np.random.RandomState(seed)
    xyz = rs.rand(h, w, 3).astype(np.float32)
    cnt = rs.randint(0, 5, size=(h, w)).astype(np.uint32)
    aov = (rs.rand(h, w, 10) * 8).astype(np.float32)
    return xyz, cnt, aov


def test_output_files_follow_the_reference_layout(tmp_path):
    scene = prb.Scene.from_string(SCENE)
    h = prb.host_lib()
    h.prh_output_file_count.restype = C.c_uint32
    h.prh_output_file_count.argtypes = [C.c_void_p]
    h.prh_save_outputs.restype = C.c_int
    h.prh_save_outputs.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    assert h.prh_output_file_count(scene._h) == 2
    W, H = scene.width, scene.height
    xyz, cnt, aov = _film(W, H)
    n = h.prh_save_outputs(scene._h, str(tmp_path).encode(), xyz.ctypes.data, cnt.ctypes.data, aov.ctypes.data, 0)
    assert n == 2
    # ---- results/image.exr: plain colour, linear sRGB from XYZ (RGBConverter::fromXYZ), clamped at 0
    attrs, chans, pl = read_exr(os.path.join(tmp_path, "results", "image.exr"))
    assert chans == ["B", "G", "R"]  # stored alphabetically
    assert struct.unpack("<4i", attrs["displayWindow"][1]) == (0, 0, W - 1, H - 1)
    assert attrs["compression"][1] == b"\0" and attrs["lineOrder"][1] == b"\0"
    M = np.array([[3.240970e+00, -1.537383e+00, -4.986108e-01], [-9.692436e-01, 1.875968e+00, 4.155506e-02], [5.563008e-02, -2.039770e-01, 1.056972e+00]], np.float32)
    rgb = np.maximum(0, xyz @ M.T)
    for k, c in enumerate("RGB"):
        np.testing.assert_allclose(pl[c], rgb[..., k], rtol=2e-6, atol=1e-7)
    # ---- results/aovs.exr
    attrs, chans, pl = read_exr(os.path.join(tmp_path, "results", "aovs.exr"))
    want = ["R", "G", "B", "[C.*L].R", "[C.*L].G", "[C.*L].B", "normal.x", "normal.y", "normal.z", "position.x", "position.y", "position.z",
            "texture.x", "texture.y", "texture.z", "normal_geometric.x", "normal_geometric.y", "normal_geometric.z", "depth", "entity_id", "sample_count"]
    assert sorted(chans) == sorted(want) and chans == sorted(chans)
    for k, c in enumerate("RGB"):  # :color 'xyz' -> the film as it is
        assert np.array_equal(pl[c], xyz[..., k])
    sf = np.where(cnt == 0, 1.0, 1.0 / np.maximum(cnt, 1)).astype(np.float32)  # technical AOVs: sums / sample count
    for k, c in enumerate(("normal.x", "normal.y", "normal.z", "position.x", "position.y", "position.z", "texture.x", "texture.y")):
        np.testing.assert_allclose(pl[c], sf * aov[..., k], rtol=1e-6)
    assert not pl["texture.z"].any()
    np.testing.assert_allclose(pl["depth"], sf * aov[..., 8], rtol=1e-6)
    np.testing.assert_allclose(pl["entity_id"], sf * aov[..., 9], rtol=1e-6)
    assert np.array_equal(pl["sample_count"], cnt.astype(np.float32))  # counters are not weighted
    for k, c in enumerate("xyz"):  # Surface.N is a copy of Geometry.N on this path (IntersectionPoint::setForSurface)
        np.testing.assert_allclose(pl["normal_geometric." + c], sf * aov[..., k], rtol=1e-6)
    assert not pl["[C.*L].R"].any()  # no expression film handed to the writer here: zeros, like a missing channel


def test_an_independent_reader_decodes_the_file(tmp_path):
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cv2 = pytest.importorskip("cv2")
    scene = prb.Scene.from_string(SCENE)
    h = prb.host_lib()
    h.prh_save_outputs.restype = C.c_int
    h.prh_save_outputs.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    xyz, cnt, _ = _film(scene.width, scene.height, seed=9)
    assert h.prh_save_outputs(scene._h, str(tmp_path).encode(), xyz.ctypes.data, cnt.ctypes.data, None, 3) == 2
    path = os.path.join(tmp_path, "results_3", "image.exr")  # context index > 0 -> results_<index>
    try:
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    except cv2.error:
        pytest.skip("OpenCV built without OpenEXR")
    if img is None:
        pytest.skip("OpenCV built without OpenEXR")
    _, _, pl = read_exr(path)
    assert img.shape == (scene.height, scene.width, 3)
    for k, c in enumerate("BGR"):
        assert np.array_equal(img[..., k], pl[c])
