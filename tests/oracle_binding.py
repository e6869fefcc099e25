"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY (the checker, never the product)."""
import ctypes as C
import os
import subprocess

import numpy as np

import pearray_b200 as prb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            build()
        l = C.CDLL(path)
        l.orc_scene_create.restype = C.c_void_p
        l.orc_scene_create.argtypes = [C.POINTER(prb.SceneDesc)]
        l.orc_scene_destroy.argtypes = [C.c_void_p]
        l.orc_render.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(prb.Tile), C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        l.orc_render_lpe.argtypes = l.orc_render.argtypes + [C.c_void_p, C.c_void_p]
        l.orc_log_fragments.restype = C.c_size_t
        l.orc_log_fragments.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t]
        l.orc_apply_filter.argtypes = [C.POINTER(prb.SceneDesc), C.c_void_p, C.c_void_p]
        l.orc_trace_closest.argtypes = [C.c_void_p, C.POINTER(prb.RaySoA), C.c_size_t, C.POINTER(prb.HitSoA), C.c_int]
        l.orc_trace_any.argtypes = [C.c_void_p, C.POINTER(prb.RaySoA), C.c_size_t, C.c_void_p, C.c_int]
        l.orc_generate_camera_rays.restype = C.c_size_t
        l.orc_generate_camera_rays.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(prb.Tile), C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_size_t]
        l.orc_material_eval.argtypes = [C.c_void_p, C.POINTER(prb.MaterialQuery), C.c_size_t, C.POINTER(prb.MaterialResult)]
        l.orc_material_sample.argtypes = [C.c_void_p, C.POINTER(prb.MaterialQuery), C.c_size_t, C.POINTER(prb.MaterialResult)]
        for n in ("orc_fresnel_dielectric", "orc_fresnel_schlick"):
            getattr(l, n).restype = C.c_float
            getattr(l, n).argtypes = [C.c_float] * 3
        l.orc_fresnel_conductor.restype = C.c_float
        l.orc_fresnel_conductor.argtypes = [C.c_float] * 4
        for n in ("orc_ndf_ggx_iso", "orc_pdf_ggx_iso"):
            getattr(l, n).restype = C.c_float
            getattr(l, n).argtypes = [C.c_void_p, C.c_float]
        for n in ("orc_ndf_ggx_aniso", "orc_pdf_ggx_aniso"):
            getattr(l, n).restype = C.c_float
            getattr(l, n).argtypes = [C.c_void_p, C.c_float, C.c_float]
        l.orc_microfacet_reflection_eval_conductor.restype = C.c_float
        l.orc_microfacet_reflection_eval_conductor.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int, C.c_float, C.c_float]
        l.orc_reflect.argtypes = [C.c_void_p] * 3
        l.orc_refract.argtypes = [C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        l.orc_halfway_reflection.argtypes = [C.c_void_p] * 3
        l.orc_cos_hemi.argtypes = [C.c_float, C.c_float, C.c_void_p]
        l.orc_tangent_frame.argtypes = [C.c_void_p] * 3
        l.orc_cartesian_from_uv.argtypes = [C.c_float, C.c_float, C.c_void_p]
        l.orc_uv_from_normal.argtypes = [C.c_void_p] * 2
        l.orc_random_stream.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]
        l.orc_eval_node.restype = C.c_float
        l.orc_eval_node.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_float]
        for n in ("orc_microfacet_reflection_eval", "orc_microfacet_reflection_pdf"):
            getattr(l, n).restype = C.c_float
            getattr(l, n).argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int]
        l.orc_halfway_refractive.argtypes = [C.c_float, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
        l.orc_sample_continuous.restype = C.c_float
        l.orc_sample_continuous.argtypes = [C.c_void_p, C.c_int, C.c_float, C.POINTER(C.c_float)]
        l.orc_sample_discrete.restype = C.c_int
        l.orc_sample_discrete.argtypes = [C.c_void_p, C.c_int, C.c_float, C.POINTER(C.c_float)]
        _lib = l
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def f3(v):
    return np.ascontiguousarray(v, dtype=np.float32)


class OracleScene:
    def __init__(self, scene):
        self.scene = scene
        self._h = lib().orc_scene_create(scene.desc)

    def render(self, tiles, first_iteration, iteration_count, rng=None, threads=None, film=None, count=None, aov=True, feedback=None, variance=None, lpe=None):
        """returns dict(film (unfiltered mean), filtered, count, aov, stats, rng, feedback[, online_mean, online_variance]);
        variance: True (fresh buffers) or a (mean, var) pair to continue"""
        w, h = self.scene.width, self.scene.height
        rng = self.scene.rng_map() if rng is None else np.array(rng, dtype=np.uint64, copy=True)
        film = np.zeros((h, w, 3), np.float32) if film is None else film
        count = np.zeros((h, w), np.uint32) if count is None else count
        aovb = np.zeros((h, w, 10), np.float32) if aov else None
        aove = np.zeros((h, w, 11), np.float32) if aov else None
        stats = np.zeros(11, np.uint64)
        feedback = np.zeros((h, w), np.uint32) if feedback is None else feedback
        threads = threads or os.cpu_count() or 1
        arr = prb.make_tiles(tiles)
        vmean = vvar = None
        if variance is True:
            vmean, vvar = np.zeros((h, w, 3), np.float32), np.zeros((h, w, 3), np.float32)
        elif variance:
            vmean, vvar = variance
        n_lpe = int(self.scene.desc.contents.n_lpe)
        lpe = np.zeros((n_lpe, h, w, 3), np.float32) if n_lpe and lpe is None else lpe
        lib().orc_render_lpe(self._h, _p(rng), arr, len(tiles), first_iteration, iteration_count, _p(film), _p(count),
                             _p(aovb) if aov else None, _p(stats), threads, _p(feedback), _p(vmean) if vmean is not None else None,
                             _p(vvar) if vvar is not None else None, _p(lpe) if n_lpe else None, _p(aove) if aov else None)
        filtered = np.empty_like(film)
        lib().orc_apply_filter(self.scene.desc, _p(film), _p(filtered))
        lpe_filtered = None
        if n_lpe:
            lpe_filtered = np.empty_like(lpe)
            for k in range(n_lpe):
                lib().orc_apply_filter(self.scene.desc, _p(lpe[k]), _p(lpe_filtered[k]))
        names = ["camera_ray_count", "light_ray_count", "primary_ray_count", "bounce_ray_count", "shadow_ray_count", "monochrome_ray_count",
                 "pixel_sample_count", "entity_hit_count", "background_hit_count", "camera_depth_count", "light_depth_count"]
        return dict(film=film, filtered=filtered, count=count, aov=aovb, stats=dict(zip(names, (int(x) for x in stats))), rng=rng, feedback=feedback,
                    online_mean=vmean, online_variance=vvar, lpe=lpe, lpe_filtered=lpe_filtered, aov_ext=aove)

    FRAG_FIELDS = [("kind", 1), ("flags", 1), ("depth", 1), ("pixel", 1), ("mis", 4), ("importance", 4), ("radiance", 4), ("pathPDF", 4),
                   ("prevPathPDF", 4), ("wvlPDF", 4), ("bsdfPDF", 4), ("lightPdfS", 1), ("roulette", 1), ("extra", 1), ("accepted", 1),
                   ("infPdfS", 4), ("wvl", 4), ("groupImportance", 4)]

    def log_fragments(self, pixels, first_iteration, iteration_count, rng=None, capacity=1 << 20):
        """every fragment the integrator pushes for the listed pixels (struct FragLog of oracle.cpp) as a dict of arrays"""
        rng = self.scene.rng_map() if rng is None else np.ascontiguousarray(rng, dtype=np.uint64)
        pixels = np.ascontiguousarray(pixels, dtype=np.uint32)
        buf = np.zeros((capacity, 48), np.float32)
        n = lib().orc_log_fragments(self._h, _p(rng), _p(pixels), pixels.size, first_iteration, iteration_count, _p(buf), capacity)
        assert n <= capacity, "fragment log overflow"
        buf = buf[:n]
        out, o = {}, 0
        for name, w in self.FRAG_FIELDS:
            out[name] = buf[:, o] if w == 1 else buf[:, o:o + w]
            o += w
        return out

    @staticmethod
    def _soa(o, d, tmin, tmax, keep):
        cols = [np.ascontiguousarray(o[:, i], dtype=np.float32) for i in range(3)] + [np.ascontiguousarray(d[:, i], dtype=np.float32) for i in range(3)]
        cols.append(None if tmin is None else np.ascontiguousarray(tmin, dtype=np.float32))
        cols.append(None if tmax is None else np.ascontiguousarray(tmax, dtype=np.float32))
        keep.extend(cols)
        return prb.RaySoA(*[None if c is None else c.ctypes.data for c in cols])

    def trace_closest(self, origins, dirs, tmin=None, tmax=None, threads=None):
        n = len(origins)
        keep = []
        rays = self._soa(np.asarray(origins), np.asarray(dirs), tmin, tmax, keep)
        ent = np.empty(n, np.uint32); prim = np.empty(n, np.uint32)
        u = np.empty(n, np.float32); v = np.empty(n, np.float32); t = np.empty(n, np.float32)
        hits = prb.HitSoA(ent.ctypes.data, prim.ctypes.data, u.ctypes.data, v.ctypes.data, t.ctypes.data)
        lib().orc_trace_closest(self._h, C.byref(rays), n, C.byref(hits), threads or os.cpu_count() or 1)
        return ent, prim, u, v, t

    def trace_any(self, origins, dirs, tmin=None, tmax=None, threads=None):
        n = len(origins)
        keep = []
        rays = self._soa(np.asarray(origins), np.asarray(dirs), tmin, tmax, keep)
        occ = np.empty(n, np.uint8)
        lib().orc_trace_any(self._h, C.byref(rays), n, _p(occ), threads or os.cpu_count() or 1)
        return occ

    def generate_camera_rays(self, tiles, iteration, rng=None):
        rng = self.scene.rng_map() if rng is None else rng
        arr = prb.make_tiles(tiles)
        cap = sum((t[2] - t[0]) * (t[3] - t[1]) for t in tiles)
        org = np.empty((cap, 3), np.float32); dr = np.empty((cap, 3), np.float32)
        wvl = np.empty((cap, 4), np.float32); pix = np.empty(cap, np.uint32)
        n = lib().orc_generate_camera_rays(self._h, _p(rng), arr, len(tiles), iteration, _p(org), _p(dr), _p(wvl), _p(pix), cap)
        return org[:n], dr[:n], wvl[:n], pix[:n]

    def material_eval(self, queries):
        out = (prb.MaterialResult * len(queries))()
        lib().orc_material_eval(self._h, queries, len(queries), out)
        return out

    def material_sample(self, queries):
        out = (prb.MaterialResult * len(queries))()
        lib().orc_material_sample(self._h, queries, len(queries), out)
        return out

    def light_sample_and_eval(self, light_id, P, wavelengths, rng_state):
        """Light::sample for a point (NEE form) + IInfiniteLight::eval in the sampled direction; returns a dict and the new RNG state"""
        f = lib().orc_light_sample_and_eval
        f.restype = None
        f.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.POINTER(C.c_float)]
        p = (C.c_float * 3)(*[float(x) for x in P])
        w = (C.c_float * 4)(*[float(x) for x in wavelengths])
        st = C.c_uint64(int(rng_state))
        out = (C.c_float * 14)()
        f(self._h, light_id, p, w, C.byref(st), out)
        o = list(out)
        return {"outgoing": np.array(o[0:3], np.float32), "pdf": o[3], "radiance": np.array(o[4:8], np.float32), "eval_pdf": o[8],
                "eval_radiance": np.array(o[9:13], np.float32), "delta": bool(o[13])}, st.value

    def eval_node(self, node, wavelength, u=0.0, v=0.0):
        return float(lib().orc_eval_node(self._h, node, wavelength, u, v))

    def close(self):
        if self._h:
            lib().orc_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
