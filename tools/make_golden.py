#!/usr/bin/env python3
"""Freeze oracle outputs as golden fixtures under tests/golden/ (nothing in the reference pins
traceRays / the integrator / the film numerically, so the oracle -- once it passes the reference's known-answer
tests, tests/test_oracle_known_answers.py -- becomes the pinned reference and its outputs are frozen here).

  python tools/make_golden.py            # rewrites tests/golden/<scene>.npz

Per scene (C0 evaluation, C1-C4, complex with the env light, the material zoo):
  rays_o/rays_d       4096 camera rays of iteration 0 (strided over the film) + 4096 incoherent rays leaving their hits
  hit_*               closest hit of every ray: entity, primitive, u, v, t          (bit exact contract)
  occ                 any-hit result of the incoherent rays with tmax = 0.75 * scene radius
  tile, film, count   unfiltered XYZ mean + sample counts of a 48x48 tile after 4 iterations (seed 42 RNG map)
  rng_after           the tile's RNG states after those 4 iterations (proves the draw ORDER, not only the image)
  stats               the 11 RenderStatistics counters of that render
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pearray_b200 as prb  # noqa: E402
from oracle_binding import OracleScene  # noqa: E402
from scene_strings import LPE_ZOO, MATERIAL_ZOO, MATERIAL_ZOO2, MATERIAL_ZOO3, MATERIAL_ZOO4, SKYSUN_ZOO  # noqa: E402

GOLDEN_ITER = 4
GOLDEN_TILE = 48


def golden_tile(scene):
    w, h = scene.width, scene.height
    n = min(GOLDEN_TILE, w, h)
    sx, sy = (w - n) // 2, (h - n) // 2 + (h // 8 if h >= 8 * n // 4 else 0)
    sy = min(sy, h - n)
    return (sx, sy, sx + n, sy + n)


def golden_rays(scene, ora):
    """deterministic ray set: strided camera rays + incoherent rays leaving the camera hits"""
    w, h = scene.width, scene.height
    org, dr, wvl, pix = ora.generate_camera_rays([(0, 0, w, h)], 0)
    step = max(1, len(org) // 4096)
    sel = np.arange(0, len(org), step)[:4096]
    o1, d1 = org[sel].copy(), dr[sel].copy()
    ent, prim, u, v, t = ora.trace_closest(o1, d1)
    hit = ent != prb.INVALID_ID
    P = (o1[hit] + d1[hit] * t[hit, None]).astype(np.float32)
    rs = np.random.RandomState(20261017)
    d2 = rs.normal(size=P.shape)
    d2 = (d2 / np.linalg.norm(d2, axis=1, keepdims=True)).astype(np.float32)
    return o1, d1, P, d2


def make(name, scene):
    ora = OracleScene(scene)
    o1, d1, o2, d2 = golden_rays(scene, ora)
    out = {}
    for tag, o, d in (("cam", o1, d1), ("inc", o2, d2)):
        ent, prim, u, v, t = ora.trace_closest(o, d)
        out.update({tag + "_o": o, tag + "_d": d, tag + "_ent": ent, tag + "_prim": prim, tag + "_u": u, tag + "_v": v, tag + "_t": t})
    from pearray_b200 import host_lib
    radius = float(host_lib().prh_scene_radius(scene._h))
    tmax = np.full(len(o2), 0.75 * radius, np.float32)
    out["inc_tmax"] = tmax
    out["inc_occ"] = ora.trace_any(o2, d2, None, tmax)
    tile = golden_tile(scene)
    r = ora.render([tile], 0, GOLDEN_ITER)
    sx, sy, ex, ey = tile
    out["tile"] = np.array(tile, np.uint32)
    out["film"] = r["film"][sy:ey, sx:ex].copy()
    out["count"] = r["count"][sy:ey, sx:ex].copy()
    out["rng_after"] = r["rng"].reshape(scene.height, scene.width)[sy:ey, sx:ex].copy()
    out["stats"] = np.array(list(r["stats"].values()), np.uint64)
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print("%-24s rays %d+%d  tile %s  film mean %s  -> %d bytes" % (name, len(o1), len(o2), tile, out["film"].mean(axis=(0, 1)), os.path.getsize(path)))


def import_reference_image():
    """examples/evaluation/cbox.exr -- the one golden image of the reference that pins this path end to end (a Mitsuba 2
    render of examples/evaluation/scene.prc == scenes/c0_evaluation.prc: 'direct' depth 6, 128 spp, linear sRGB).  It is
    frozen as 8x8 block means (32x32x3 float32) plus the mask of blocks that see the luminaire directly; a cross-renderer
    image is a sanity bound, not a bit-level one."""
    ref = "/root/reference/examples/evaluation/cbox.exr"
    if not os.path.exists(ref):
        print("reference checkout not mounted: keeping tests/golden/cbox_reference_blocks.npz")
        return
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    import cv2
    img = cv2.imread(ref, cv2.IMREAD_UNCHANGED)[..., ::-1].astype(np.float32)  # BGR -> RGB
    lum = (img.max(axis=2) > 2.0).astype(np.uint8)
    lum = cv2.dilate(lum, np.ones((9, 9), np.uint8)).astype(bool)
    w = (~lum).astype(np.float32)[..., None]
    blocks = (img * w).reshape(32, 8, 32, 8, 3).sum(axis=(1, 3)) / np.maximum(w.reshape(32, 8, 32, 8, 1).sum(axis=(1, 3)), 1)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cbox_reference_blocks.npz"), blocks=blocks.astype(np.float32), pixel_mask=~lum)
    print("cbox_reference_blocks: mean rgb", blocks.mean(axis=(0, 1)))


if __name__ == "__main__":
    only = set(sys.argv[1:])  # python tools/make_golden.py [name ...]: only the named fixtures
    want = lambda n: not only or n in only
    if want("cbox_reference_blocks"):
        import_reference_image()
    for n in ("c0_evaluation", "c1_sphere", "c2_cornellbox", "c3_cornellbox_glassy", "c4_boltsandgears", "c4b_complex_env", "c4c_complex"):
        if want(n):
            make(n, prb.Scene.from_file(os.path.join(ROOT, "scenes", n + ".prc")))
    for n, src in (("material_zoo", MATERIAL_ZOO), ("skysun_zoo", SKYSUN_ZOO), ("material_zoo2", MATERIAL_ZOO2), ("material_zoo3", MATERIAL_ZOO3),
                   ("material_zoo4", MATERIAL_ZOO4), ("lpe_zoo", LPE_ZOO)):
        if want(n):
            make(n, prb.Scene.from_string(src))
