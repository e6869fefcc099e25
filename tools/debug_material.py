#!/usr/bin/env python3
"""Print the material-sample queries where GPU and oracle differ (development tool; runs on the GPU box).
  python tools/debug_material.py scene.prc material_id"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pearray_b200 as prb  # noqa: E402
from oracle_binding import OracleScene  # noqa: E402

path, mat = sys.argv[1], int(sys.argv[2])
scene = prb.Scene.from_file(path)
ctx = prb.Context(0)
ctx.upload_scene(scene)
ora = OracleScene(scene)
rs = np.random.RandomState(7)
n = 2048
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
g = ctx.material_sample(q)
o = ora.material_sample(q)
bad = 0
for i in range(n):
    ga = np.array([*g[i].weight, *g[i].pdf_s, *g[i].L], dtype=np.float32)
    oa = np.array([*o[i].weight, *o[i].pdf_s, *o[i].L], dtype=np.float32)
    if not np.array_equal(ga.view(np.uint32), oa.view(np.uint32)) or g[i].type != o[i].type:
        bad += 1
        if bad <= 12:
            print("q%d V %s wvl %s" % (i, [float(np.float32(x)).hex() for x in q[i].V], list(q[i].wavelength_nm)))
            print("   gpu type %d w %s pdf %s L %s" % (g[i].type, list(g[i].weight), list(g[i].pdf_s), [float(x).hex() for x in g[i].L]))
            print("   ora type %d w %s pdf %s L %s" % (o[i].type, list(o[i].weight), list(o[i].pdf_s), [float(x).hex() for x in o[i].L]))
print("mismatching queries: %d of %d" % (bad, n))
