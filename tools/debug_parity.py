#!/usr/bin/env python3
"""GPU-vs-oracle bisection helper (development tool; runs on the GPU box).

  python tools/debug_parity.py [scene.prc]

Renders one tile with different integrator settings (depth 1, NEE off, ...) on the GPU and with the oracle and
prints where the two start to disagree."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pearray_b200 as prb  # noqa: E402
from oracle_binding import OracleScene  # noqa: E402


def compare(name, path, tile, iters, **over):
    scene = prb.Scene.from_file(path)
    s = scene.settings
    for k, v in over.items():
        setattr(s, k, v)
    ctx = prb.Context(0)
    ctx.upload_scene(scene)
    rng = scene.rng_map()
    ctx.upload_rng(rng)
    ctx.render_tiles([tile], 0, iters)
    xyz, cnt = ctx.film()
    ora = OracleScene(scene)
    ref = ora.render([tile], 0, iters, rng=rng)
    sx, sy, ex, ey = tile
    a = xyz[sy:ey, sx:ex].astype(np.float64)
    b = ref["filtered"][sy:ey, sx:ex].astype(np.float64)
    rel = np.sqrt(np.mean((a - b) ** 2)) / max(1e-12, np.mean(b))
    bad = np.abs(a - b).max(axis=2) > 1e-4 * (np.abs(b).max(axis=2) + 1e-3)
    st = ctx.stats().as_dict()
    print("== %s %s: relRMSE %.3g, pixels differing %.4f, mean gpu %s oracle %s" % (name, over, rel, bad.mean(), a.mean(axis=(0, 1)), b.mean(axis=(0, 1))))
    for k, v in ref["stats"].items():
        if st[k] != v:
            print("   stat %-22s gpu %10d oracle %10d" % (k, st[k], v))
    grng = ctx.download_rng()
    m = cnt.reshape(-1) >= 0
    same_rng = np.mean(grng == ref["rng"])
    print("   rng states equal: %.4f   sample counts equal: %s" % (same_rng, np.array_equal(cnt, ref["count"])))
    ys, xs = np.nonzero(bad)
    for i in range(min(4, len(ys))):
        print("   px (%d,%d) gpu %s oracle %s" % (xs[i] + sx, ys[i] + sy, a[ys[i], xs[i]], b[ys[i], xs[i]]))
    ctx.close()
    return rel


def materials(path):
    scene = prb.Scene.from_file(path)
    ctx = prb.Context(0)
    ctx.upload_scene(scene)
    ora = OracleScene(scene)
    nm = scene.desc.contents.n_materials
    rs = np.random.RandomState(7)
    n = 256
    for mat in range(nm):
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
            exact = np.mean(ga.view(np.uint32) == oa.view(np.uint32))
            err = np.nanmax(np.abs(ga - oa) / (np.abs(oa) + 1e-6))
            print("   material %d type %d %-6s: bit-exact fields %.4f, max rel err %.3g, flags/type/rng equal %s" %
                  (mat, scene.desc.contents.materials[mat].type, kind, exact, err, np.array_equal(gf, of)))
    ctx.close()


def shadows(path):
    scene = prb.Scene.from_file(path)
    ctx = prb.Context(0)
    ctx.upload_scene(scene)
    ctx.upload_rng(scene.rng_map())
    ora = OracleScene(scene)
    org, dr, wvl, pix = ctx.generate_camera_rays([(0, 0, scene.width, scene.height)], 0)
    ent, prim, u, v, t = ctx.trace_closest(org, dr)
    hit = ent != prb.INVALID_ID
    P = org[hit] + dr[hit] * t[hit, None]
    rs = np.random.RandomState(3)
    tgt = rs.uniform(-1, 1, size=P.shape).astype(np.float32) * float(np.abs(P).max())
    d = tgt - P
    dist = np.linalg.norm(d, axis=1).astype(np.float32)
    d = (d / dist[:, None]).astype(np.float32)
    tmin = np.full(len(P), 1e-4, np.float32)
    tmax = (dist - 1e-3).astype(np.float32)
    g = ctx.trace_any(P, d, tmin, tmax)
    o = ora.trace_any(P, d, tmin, tmax)
    print("   any-hit rays %d: occluded gpu %d oracle %d, mismatches %d" % (len(P), g.sum(), o.sum(), (g != o).sum()))
    g2 = ctx.trace_closest(P, d, tmin, tmax)
    o2 = ora.trace_closest(P, d, tmin, tmax)
    print("   incoherent closest: entity mismatches %d prim mismatches %d t mismatches %d" %
          ((g2[0] != o2[0]).sum(), (g2[1] != o2[1]).sum(), (g2[4].view(np.uint32) != o2[4].view(np.uint32)).sum()))
    ctx.close()


if __name__ == "__main__":
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "scenes", "c2_cornellbox.prc")
    print("scene", path)
    materials(path)
    shadows(path)
    scene = prb.Scene.from_file(path)
    w, h = scene.width, scene.height
    tile = (w // 4, h // 4, w // 4 + 64, h // 4 + 64)
    compare("depth1-no-nee", path, tile, 1, max_ray_depth=1, do_nee=0)
    compare("depth1", path, tile, 1, max_ray_depth=1)
    compare("depth2-no-nee", path, tile, 1, max_ray_depth=2, do_nee=0)
    compare("depth2", path, tile, 1, max_ray_depth=2)
    compare("depth3-no-nee", path, tile, 1, max_ray_depth=3, do_nee=0)
    compare("depth3", path, tile, 1, max_ray_depth=3)
    compare("depth4", path, tile, 1, max_ray_depth=4)
    compare("depth5", path, tile, 1, max_ray_depth=5)
    compare("depth6-soft64", path, tile, 1, max_ray_depth=6, soft_max_ray_depth=64)
    compare("depth16-soft64", path, tile, 1, max_ray_depth=16, soft_max_ray_depth=64)
    compare("depth6-soft2", path, tile, 1, max_ray_depth=6, soft_max_ray_depth=2)
    compare("full-1it", path, tile, 1)
    compare("full-4it", path, tile, 4)
