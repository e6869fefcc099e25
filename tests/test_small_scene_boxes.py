"""The premise of k_trace_small's box phase (pearray_b200/csrc/dev_bvh.cuh: traverseSmall, prb_api.cu: the face boxes), checked on the
CPU: for every ray the oracle's exhaustive triangle loop reports a hit for, the padded world-space box of the hit triangle -- built
with the host's formula (float64 transform of the local-space vertices, pad 8e-6 of the largest coordinate, one ulp outwards) -- is
met by the ray under the device's slab test (one fused multiply-add per plane, reciprocal direction, plane parameters widened by
4e-6 relative), with the hit distance inside the clipped interval.  A face box is the union of its triangles' boxes with at least
their pad, so the face test passes whenever the triangle's own box does.  The GPU suite proves the same end to end (bit-identical
films); this test sweeps far more rays, grazing ones included, without a GPU."""
import ctypes as C

import numpy as np
import pytest

import pearray_b200 as prb
from oracle_binding import OracleScene

T_SLACK = np.float32(4e-6)  # SMALL_T_SLACK


def _triangles(scene):
    """(entity, prim) -> list of world-space triangles (3x3 float64) of the small-scene list, by the walk prb_upload_scene does"""
    d = scene.desc.contents
    tri_raw = np.frombuffer((C.c_char * (int(d.n_bvh_tris) * 48)).from_address(d.bvh_tris), dtype=np.uint8).reshape(-1, 48)
    node_raw = np.frombuffer((C.c_char * (int(d.n_bvh_nodes) * 80)).from_address(d.bvh_nodes), dtype=np.uint8).reshape(-1, 80)
    out = {}
    for e in range(d.n_entities):
        en = d.entities[e]
        if en.type == 1:  # sphere: analytic
            continue
        m = np.array(list(en.local_to_world), np.float64).reshape(3, 4)
        todo = [int(en.blas_root)]
        while todo:
            n = node_raw[todo.pop()]
            child_base, prim_base = int(n[16:20].view(np.uint32)[0]), int(n[20:24].view(np.uint32)[0])
            for meta in n[24:32]:
                meta = int(meta)
                if meta == 0xFF:
                    continue
                if meta & 0x80:
                    todo.append(child_base + (meta & 0x7F))
                    continue
                for k in range(((meta >> 5) & 3) + 1):
                    t = tri_raw[prim_base + (meta & 0x1F) + k]
                    v = np.stack([t[0:12].view(np.float32), t[16:28].view(np.float32), t[32:44].view(np.float32)]).astype(np.float64)
                    if en.type == 0:  # mesh: local space -> world
                        v = v @ m[:, :3].T + m[:, 3]
                    out.setdefault((e, int(t[12:16].view(np.uint32)[0])), []).append(v)
    return out


def _box(v):
    lo, hi = v.min(axis=0), v.max(axis=0)
    pad = max(8e-6 * np.abs(v).max(), 1e-30)
    return (np.nextafter((lo - pad).astype(np.float32), np.float32(-np.inf)), np.nextafter((hi + pad).astype(np.float32), np.float32(np.inf)))


def _slab(lo, hi, o, d, tmin, tmax):
    """smallBoxMask for one box: fp32, every plane parameter one fused multiply-add"""
    a = np.where(np.abs(d) > 1e-20, d, np.copysign(np.float32(1e-20), d)).astype(np.float32)
    inv = (np.float32(1) / a).astype(np.float32)  # (the device's rcp.approx is within one ulp of this)
    oi = (-(o * inv)).astype(np.float32)
    fma = lambda x, y, z: (x.astype(np.float64) * y.astype(np.float64) + z.astype(np.float64)).astype(np.float32)
    ta, tc = fma(lo, inv, oi), fma(hi, inv, oi)
    tn, tf = np.minimum(ta, tc).max(), np.maximum(ta, tc).min()
    tn = np.float32(tn - abs(tn) * T_SLACK)
    tf = np.float32(tf + abs(tf) * T_SLACK)
    return max(tn, np.float32(tmin)), min(tf, np.float32(tmax))


def _check(scene, rng_seed, n=6000):
    tris = _triangles(scene)
    rs = np.random.RandomState(rng_seed)
    allv = np.concatenate([v for vs in tris.values() for v in vs])
    lo, hi = allv.min(axis=0), allv.max(axis=0)
    # rays from inside the scene's bounds towards random triangle points (a third of them aimed at an edge or a vertex: grazing hits)
    keys = list(tris)
    org = rs.uniform(lo + 0.02 * (hi - lo), hi - 0.02 * (hi - lo), size=(n, 3)).astype(np.float32)
    tgt = np.empty((n, 3))
    for i in range(n):
        vs = tris[keys[rs.randint(len(keys))]]
        v = vs[rs.randint(len(vs))]
        w = rs.dirichlet([1, 1, 1])
        if i % 3 == 0:
            w[rs.randint(3)] = 0.0
            w /= w.sum()
        tgt[i] = w @ v
    dr = (tgt - org).astype(np.float32)
    dr /= np.maximum(np.linalg.norm(dr, axis=1, keepdims=True), 1e-12)
    tmin = np.full(n, 1e-4, np.float32)
    ent, prim, u, v, t = OracleScene(scene).trace_closest(org, dr, tmin=tmin)
    hit = ent != 0xFFFFFFFF
    assert hit.sum() > n // 2
    culled = 0
    for i in np.nonzero(hit)[0]:
        if (int(ent[i]), int(prim[i])) not in tris:  # an analytic sphere
            continue
        ok = False
        for tv in tris[(int(ent[i]), int(prim[i]))]:
            blo, bhi = _box(tv)
            t0, t1 = _slab(blo, bhi, org[i], dr[i], 1e-4, np.inf)
            ok = ok or (t0 <= t1 and t0 <= t[i] <= t1)
        culled += not ok
    assert culled == 0


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_face_boxes_never_cull_a_hit(seed):
    from test_gpu_parity import _many_faces_scene
    _check(prb.Scene.from_string(_many_faces_scene(seed)), 100 + seed)


@pytest.mark.parametrize("name", ["c2_cornellbox", "c3_cornellbox_glassy", "c0_evaluation"])
def test_face_boxes_never_cull_a_hit_on_the_config_scenes(name):
    from conftest import scene_path
    _check(prb.Scene.from_file(scene_path(name + ".prc")), 7)
