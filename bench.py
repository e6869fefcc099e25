#!/usr/bin/env python3
"""bench.py -- spectral path samples/s of the 'direct' (unidirectional PT) hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scene c2|c1|c3|c4|c4b|c4c|c0|c5] [--spp S]
                  [--partition samples|tiles]

Workload (BASELINE.json configs[1]): cornellbox.prc, 500x500, 'direct' integrator depth 6, mjitt sampler,
1024 spp, diffuse materials + one area light.  One "step" = one complete render of that configuration
(500*500*1024 = 262 144 000 spectral path samples, 4 wavelengths each) through prb_render_tiles.

Our arm prints one JSON line:
  value   samples/s with the scene, RNG map and film resident in HBM (CUDA events via the C ABI, max over ranks)
  e2e     samples/s through the C ABI with HOST buffers: per step the RNG map is uploaded, the tiles rendered and the
          film (xyz + sample counts) downloaded inside the timed region
  roofline / cpu_baseline / clocks / gpu_launches as the contract asks (see DESIGN.md section "Measurement").
Multi-GPU (torchrun, one rank per GPU): scene replicated, work partitioned by sample ranges (default; every rank
renders the full film for its own 1024-iteration range with a decorrelated RNG map -> weak scaling) or by interleaved
tiles (--partition tiles; bit-identical to 1 GPU, strong scaling); films combined by one NCCL reduce to rank 0
inside the timed region.

--impl reference: the reference's own CPU implementation cannot be built here (Eigen/Embree/TBB/OIIO absent, SURVEY
F4), so the arm times oracle/ (the CPU restatement, std::thread x all host cores) on a bounded sample of the same
workload.  This file and tests/ are the only places that execute oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SCENES = {
    "c1": ("c1_sphere.prc", "sphere.prc 1000x1000 direct sobol 64spp"),
    "c2": ("c2_cornellbox.prc", "cornellbox.prc 500x500 direct(depth 6) mjitt 1024spp"),
    "c3": ("c3_cornellbox_glassy.prc", "cornellbox_glassy.prc 256x256 direct(depth 16, power MIS) mjitt 128spp"),
    "c4": ("c4_boltsandgears.prc", "boltsandgears.prc 1000x1000 direct mjitt 256spp"),
    "c4b": ("c4b_complex_env.prc", "complex.prc 1920x1080 direct sobol 4096spp (sky+sun replaced by a D65 env light)"),
    "c4c": ("c4c_complex.prc", "complex.prc 1920x1080 direct sobol 4096spp (as shipped: Hosek-Wilkie sky + sun)"),
    "c0": ("c0_evaluation.prc", "evaluation/scene.prc 256x256 direct(depth 6) sobol 128spp"),
}


# ----------------------------------------------------------------------------------------------- helpers
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self._stop.is_set():
                    break
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self._stop.set()
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(scene_key, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu --set full capture of this
    workload (profiles/traffic.json, written by hand from profiles/r01_ncu_*.txt); None when no capture is on file"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)[scene_key][kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def bvh_depth_bytes(n_tris):
    """minimal-path bytes of one closest-hit / any-hit ray, SURVEY 8(d): 32 + (24|4) + 80*ceil(log8(N/4)) + 4*48"""
    import math
    levels = max(1, math.ceil(math.log(max(n_tris / 4.0, 1.0001), 8)))
    return 32 + 24 + 80 * levels + 4 * 48, 32 + 4 + 80 * levels + 4 * 48


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    import numpy as np
    import pearray_b200 as prb
    from oracle_binding import OracleScene
    fname, label = SCENES[args.scene]
    scene = prb.Scene.from_file(os.path.join(ROOT, "scenes", fname))
    ora = OracleScene(scene)
    cores = os.cpu_count() or 1
    tiles = scene.tiles(8, 8)
    npix = scene.width * scene.height
    rng = scene.rng_map()
    # calibrate: 1 spp, then size a step to ~4 s of CPU work
    t0 = time.perf_counter()
    r = ora.render(tiles, 0, 1, rng=rng, threads=cores, aov=False)
    t1 = time.perf_counter() - t0
    spp = max(1, min(int(scene.settings.max_sample_count), int(round(4.0 / max(t1, 1e-3)))))
    film = np.zeros((scene.height, scene.width, 3), np.float32)
    cnt = np.zeros((scene.height, scene.width), np.uint32)
    for _ in range(args.warmup):
        ora.render(tiles, 0, 1, rng=rng, threads=cores, film=film, count=cnt, aov=False)
    t0 = time.perf_counter()
    rays = 0
    for k in range(args.steps):
        r = ora.render(tiles, k * spp, spp, rng=rng, threads=cores, film=film, count=cnt, aov=False)
        rays += r["stats"]["primary_ray_count"] + r["stats"]["bounce_ray_count"] + r["stats"]["shadow_ray_count"]
    dt = time.perf_counter() - t0
    value = npix * spp * args.steps / dt
    sample = "%d spp per step over the full %dx%d film (of %d spp), %d threads" % (spp, scene.width, scene.height, scene.settings.max_sample_count, cores)
    line = {"impl": "reference", "metric": "spectral path samples/s", "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "mrays_per_s": rays / dt / 1e6,
            "config": {"workload": label, "scene": fname, "note": "reference not buildable here (Eigen/Embree/TBB/OIIO absent): CPU oracle port, bounded sample"},
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import pearray_b200 as prb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback (use --impl reference for the CPU arm)")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    fname, label = SCENES[args.scene]
    scene = prb.Scene.from_file(os.path.join(ROOT, "scenes", fname))
    W, H = scene.width, scene.height
    spp = int(scene.settings.max_sample_count) if args.spp is None else args.spp
    all_tiles = scene.tiles(8, 8)
    from pearray_b200 import multigpu
    if args.partition == "tiles" and world > 1:
        tiles = multigpu.partition_tiles(all_tiles, rank, world)  # interleaved tiles, bit-identical for any G
        first_iter = 0
        scaling = "strong"
        seed_rank = 0
    else:
        tiles = all_tiles
        first_iter = 0  # sample-range partition: every rank renders its own spp iterations of the full film, decorrelated RNG map
        scaling = "weak"
        seed_rank = rank
    npix_rank = sum((t[2] - t[0]) * (t[3] - t[1]) for t in tiles)
    samples_rank = npix_rank * spp
    scene.settings.seed = multigpu.rank_seed(scene.settings.seed, seed_rank)  # decorrelated per-rank RNG map in sample-range mode
    ctx = prb.Context(local_rank)
    ctx.upload_scene(scene)
    rng_host = scene.rng_map()
    # pinned host buffers for the e2e leg
    rng_pinned = torch.empty(W * H, dtype=torch.int64).pin_memory()
    rng_pinned.numpy().view(np.uint64)[:] = rng_host
    film_pinned = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
    cnt_pinned = torch.empty((H, W), dtype=torch.int32).pin_memory()
    film_dev = torch.zeros((H * W * 4,), dtype=torch.float32, device=dev)  # export buffer handed to the NCCL reduce

    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)

    def reduce_films():
        """film export (own kernel) + one NCCL reduce to rank 0; returns device ms (CUDA events on torch's stream)"""
        if world == 1:
            return 0.0
        ev0.record()
        ctx.film_export_device(film_dev.data_ptr())
        multigpu.reduce_film(film_dev.view(-1, 4), "samples" if scaling == "weak" else "tiles", world)
        ev1.record()
        ev1.synchronize()
        return ev0.elapsed_time(ev1)

    split_ms = [0.0, 0.0]  # render, film reduce (the reduce of a rank that finished early includes waiting for the slowest rank)

    def step_resident():
        ctx.render_tiles(tiles, first_iter, spp)  # returns when the wavefront retired every sample
        r = ctx.last_device_ms()
        f = reduce_films()
        split_ms[0] += r
        split_ms[1] += f
        return r + f

    def step_e2e():
        t0 = time.perf_counter()
        ctx.upload_rng(rng_pinned.numpy().view(np.uint64))
        ctx.render_tiles(tiles, first_iter, spp)
        reduce_films()
        if world > 1 and rank == 0:
            ctx.film_import_device(film_dev.data_ptr())
        if rank == 0 or world == 1:
            ctx.film(out=film_pinned.numpy(), count_out=cnt_pinned.numpy().view(np.uint32))
        return 1e3 * (time.perf_counter() - t0)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx.upload_rng(rng_host)
    for _ in range(args.warmup):
        step_resident()
    ctx.reset_stats()
    split_ms[0] = split_ms[1] = 0.0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- timed: K steps, device time (CUDA events on the context stream inside prb_render_tiles), max over ranks
    barrier()
    wall0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        dev_ms += step_resident()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    dev_ms = max_over_ranks(dev_ms)
    rank_ms = None
    if world > 1:  # per-rank split of the timed region, for the scaling analysis
        t = torch.tensor(split_ms, dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        rank_ms = [{"render": float(x[0]) / args.steps, "film_reduce": float(x[1]) / args.steps} for x in allt]
    st = ctx.stats()
    launches = int(st.kernel_launches)
    rays_rank = st.ray_count
    # ---- e2e leg: host buffers through the C ABI
    for _ in range(min(args.warmup, 1)):
        step_e2e()
    barrier()
    e2e_ms = 0.0
    for _ in range(args.steps):
        e2e_ms += step_e2e()
    barrier()
    e2e_ms = max_over_ranks(e2e_ms)
    if rank == 0:
        sampler.stop()
    # ---- per-stage profile pass (rank 0): event pair around every kernel launch of one more identical step
    roofline = None
    stage = None
    if rank == 0:
        ctx.set_profiling(True)
        ctx.reset_stats()
        prof_spp = min(spp, 64)
        ctx.render_tiles(tiles, first_iter, prof_spp)
        stage = ctx.stage_times()
        pst = ctx.stats()
        ctx.set_profiling(False)
        total = sum(v[0] for v in stage.values()) or 1.0
        dom = max(stage, key=lambda k: stage[k][0])
        n_tris = int(scene.desc.contents.n_bvh_tris)
        b_closest, b_any = bvh_depth_bytes(n_tris)
        # algorithmic bytes per launch (DESIGN.md "Measurement"): trace = B_ray per closest-hit ray + B_any per any-hit ray
        # (SURVEY 8(d)); shade = wavefront state read + written per path vertex (ray 2x32, hit 20, path state 5x16 r+w,
        # RNG 16, film/accumulator 32, shadow ray out 48)
        n_closest = int(pst.primary_ray_count + pst.bounce_ray_count)
        n_any = int(pst.shadow_ray_count)
        units = {"trace": n_closest + n_any, "shade": n_closest}
        bytes_total = {"trace": n_closest * b_closest + n_any * b_any, "shade": n_closest * (64 + 20 + 160 + 16 + 32 + 48)}
        per_unit = {k: bytes_total[k] / max(units[k], 1) for k in units}
        ms_dom, n_dom = stage[dom]
        peak, peak_src = measured_peak_gbs()
        achieved = bytes_total[dom] / (ms_dom * 1e-3) / 1e9 if ms_dom > 0 else 0.0
        roofline = {"bound": "hbm", "kernel": "k_" + dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic(args.scene, "k_" + dom), "peak_source": peak_src, "avg_launch_us": 1e3 * ms_dom / max(n_dom, 1), "launches": n_dom,
                    "bytes_per_unit": per_unit[dom], "units_per_launch": units[dom] / max(n_dom, 1),
                    "stage_share": {k: v[0] / total for k, v in stage.items()},
                    "note": "scene is L2-resident (%d triangles): HBM is not the binding resource, see DESIGN.md" % n_tris}
    # ---- cpu baseline (rank 0, N=1 only): the oracle port on all host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle_binding import OracleScene
        ora = OracleScene(scene)
        cores = os.cpu_count() or 1
        t0 = time.perf_counter()
        ora.render(all_tiles, 0, 1, rng=rng_host, threads=cores, aov=False)
        t1 = time.perf_counter() - t0
        n = max(1, min(spp, int(round(12.0 / max(t1, 1e-3)))))
        t0 = time.perf_counter()
        ora.render(all_tiles, 0, n, rng=rng_host, threads=cores, aov=False)
        dt = time.perf_counter() - t0
        cpu = {"value": W * H * n / dt, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": "%d of %d spp over the full %dx%d film, oracle (C++17 restatement, std::thread x %d)" % (n, spp, W, H, cores)}
    if rank == 0:
        total_samples = samples_rank * world * args.steps if scaling == "weak" else W * H * spp * args.steps
        value = total_samples / (dev_ms * 1e-3)
        e2e = total_samples / (e2e_ms * 1e-3)
        h2d = W * H * 8 + len(tiles) * 16
        d2h = W * H * 3 * 4 + W * H * 4
        line = {"metric": "spectral path samples/s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "mrays_per_s": rays_rank * world / (dev_ms * 1e-3) / 1e6 if scaling == "weak" else None,
                "wall_ms_per_step": wall_ms / args.steps,
                "config": {"workload": label, "scene": fname, "spp_per_step": spp, "film": [W, H], "partition": args.partition if world > 1 else "single",
                           "l2": "wavefront state (%d paths x ~250 B) and film are re-written every wavefront iteration; scene is L2-resident by nature" % npix_rank},
                "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu,
                "stage_ms": {k: v[0] for k, v in stage.items()} if stage else None, "rank_ms_per_step": rank_ms}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- C5: triangle soup
def run_soup(args, rank, world, local_rank):
    """synthetic N-triangle soup (default 10 M), 2048x2048 pinhole, 16 jittered passes; three ray classes
    timed separately through prb_trace_closest_device / prb_trace_any_device with the ray streams resident in HBM:
    primary (coherent closest hit), shadow (any hit towards a point light at (0,3,0), tfar = dist - 1e-3) and incoherent
    (cosine-hemisphere bounce from every primary hit, closest hit).  The only configuration that streams the BVH from HBM."""
    import numpy as np
    import torch
    import pearray_b200 as prb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    res = args.soup_film
    t0 = time.perf_counter()
    scene = prb.Scene.soup(args.triangles, seed=1234, film=(res, res))
    build_s = time.perf_counter() - t0
    d = scene.desc.contents
    ctx = prb.Context(local_rank)
    ctx.upload_scene(scene)
    ctx.upload_rng(scene.rng_map())
    verts = torch.from_numpy(np.ctypeslib.as_array(d.vertices, shape=(d.n_vertices, 3)).copy()).to(dev).view(-1, 3, 3)
    n = res * res
    passes = args.passes
    # weak scaling over ranks: every rank traces its own passes (pass index offset by rank) against the replicated scene
    f32 = dict(dtype=torch.float32, device=dev)
    ent = torch.empty(n, dtype=torch.int32, device=dev); prim = torch.empty_like(ent)
    u = torch.empty(n, **f32); v = torch.empty(n, **f32); t = torch.empty(n, **f32)
    occ = torch.empty(n, dtype=torch.uint8, device=dev)
    light = torch.tensor([0.0, 3.0, 0.0], **f32)

    def soa(x):  # (n,3) -> three contiguous columns
        return [x[:, i].contiguous() for i in range(3)]

    def ptrs(o, dd, tmin=None, tmax=None):
        return [c.data_ptr() for c in o] + [c.data_ptr() for c in dd] + [tmin.data_ptr() if tmin is not None else None, tmax.data_ptr() if tmax is not None else None]

    hitp = [ent.data_ptr(), prim.data_ptr(), u.data_ptr(), v.data_ptr(), t.data_ptr()]
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    ms = {"primary": 0.0, "shadow": 0.0, "incoherent": 0.0}
    rays = {"primary": 0, "shadow": 0, "incoherent": 0}
    hits = {"primary": 0, "shadow": 0, "incoherent": 0}
    sampler = ClockSampler(local_rank)

    def one_pass(p, timed):
        org, dr, _, _ = ctx.generate_camera_rays([(0, 0, res, res)], p)
        o = soa(torch.from_numpy(org).to(dev)); dd = soa(torch.from_numpy(dr).to(dev))
        torch.cuda.synchronize(dev)
        ctx.trace_closest_device(ptrs(o, dd), n, hitp)
        if timed:
            ms["primary"] += ctx.last_device_ms(); rays["primary"] += n
        hit = ent != -1
        idx = hit.nonzero().squeeze(1)
        m = int(idx.numel())
        if timed:
            hits["primary"] += m
        O = torch.stack(o, 1)[idx]; D = torch.stack(dd, 1)[idx]
        P = O + D * t[idx, None]
        tri = verts[prim[idx].long()]
        N = torch.linalg.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
        N = N / N.norm(dim=1, keepdim=True).clamp_min(1e-30)
        N = torch.where((N * D).sum(1, keepdim=True) > 0, -N, N)  # face the incoming ray
        # shadow rays (Scene::traceShadowRay semantics: tnear 1e-4, tfar = distance - 1e-3)
        L = light - P
        dist_l = L.norm(dim=1)
        L = L / dist_l[:, None]
        so = soa(P); sd = soa(L)
        tmin = torch.full((m,), 1e-4, **f32); tmax = (dist_l - 1e-3).contiguous()
        torch.cuda.synchronize(dev)
        ctx.trace_any_device(ptrs(so, sd, tmin, tmax), m, occ.data_ptr())
        if timed:
            ms["shadow"] += ctx.last_device_ms(); rays["shadow"] += m; hits["shadow"] += int(occ[:m].sum())
        # incoherent: cosine-hemisphere bounce around N
        r1 = torch.rand(m, generator=gen, **f32); r2 = torch.rand(m, generator=gen, **f32)
        ct = r1.sqrt(); st_ = (1 - r1).clamp_min(0).sqrt(); ph = 2 * np.pi * r2
        a = torch.where(N[:, 0:1].abs() > 0.9, torch.tensor([0.0, 1.0, 0.0], **f32), torch.tensor([1.0, 0.0, 0.0], **f32)).expand(m, 3)
        T = torch.linalg.cross(N, a); T = T / T.norm(dim=1, keepdim=True); B = torch.linalg.cross(N, T)
        W = T * (st_ * ph.cos())[:, None] + B * (st_ * ph.sin())[:, None] + N * ct[:, None]
        bo = soa(P); bd = soa(W)
        torch.cuda.synchronize(dev)
        ctx.trace_closest_device(ptrs(bo, bd, tmin, None), m, hitp)
        if timed:
            ms["incoherent"] += ctx.last_device_ms(); rays["incoherent"] += m; hits["incoherent"] += int((ent[:m] != -1).sum())

    for w in range(max(args.warmup, 1)):
        one_pass(1000 + w, False)
    ctx.reset_stats()
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        for p in range(passes):
            one_pass(rank * passes + p, True)
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - wall0
    if rank == 0:
        sampler.stop()
    tot_ms = sum(ms.values())
    tot_rays = sum(rays.values())
    if world > 1:
        tt = torch.tensor([tot_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tot_ms_max = float(tt.item())
        rr = torch.tensor([float(tot_rays)], dtype=torch.float64, device=dev)
        dist.all_reduce(rr, op=dist.ReduceOp.SUM)
        all_rays = float(rr.item())
    else:
        tot_ms_max, all_rays = tot_ms, float(tot_rays)
    if rank == 0:
        n_tris = int(d.n_bvh_tris)
        b_closest, b_any = bvh_depth_bytes(n_tris)
        peak, peak_src = measured_peak_gbs()
        per_class = {}
        for k in ms:
            bpr = b_any if k == "shadow" else b_closest
            per_class[k] = {"mrays_per_s": rays[k] / (ms[k] * 1e-3) / 1e6, "hit_fraction": hits[k] / max(rays[k], 1), "ms": ms[k],
                            "achieved_gbs": rays[k] * bpr / (ms[k] * 1e-3) / 1e9, "frac": rays[k] * bpr / (ms[k] * 1e-3) / 1e9 / peak, "bytes_per_ray": bpr}
        dom = max(ms, key=lambda k: ms[k])
        launches = args.steps * passes
        roofline = {"bound": "hbm", "kernel": "k_trace_any" if dom == "shadow" else "k_trace_closest (%s rays)" % dom, "achieved": per_class[dom]["achieved_gbs"],
                    "peak": peak, "unit": "GB/s", "frac": per_class[dom]["frac"], "traffic": ncu_traffic("c5", "k_trace_any" if dom == "shadow" else "k_trace_closest"), "peak_source": peak_src,
                    "avg_launch_us": 1e3 * ms[dom] / launches, "launches": launches, "bytes_per_unit": per_class[dom]["bytes_per_ray"],
                    "units_per_launch": rays[dom] / launches}
        line = {"metric": "rays/s (primary+shadow+incoherent)", "value": all_rays / (tot_ms_max * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": tot_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": "synthetic %d-triangle soup, %dx%d pinhole, %d passes/step, primary+shadow+1-bounce incoherent" % (n_tris, res, res, passes),
                                                "bvh_nodes": int(d.n_bvh_nodes), "bvh_mbytes": (int(d.n_bvh_nodes) * 80 + n_tris * 48) / 1e6, "host_build_s": build_s,
                                                "l2": "BVH + triangles (%.0f MB) exceed the 126 MB L2 for >= 2.4 M triangles" % ((int(d.n_bvh_nodes) * 80 + n_tris * 48) / 1e6)},
                "classes": per_class, "roofline": roofline, "gpu_launches": int(ctx.stats().kernel_launches), "clocks": sampler.summary(),
                "e2e": None, "cpu_baseline": None, "wall_s": wall}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_JSON_FD = None


def emit(line):
    """the ONE JSON line goes to the real stdout; everything else (NCCL banners, library chatter) was moved to stderr"""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)  # NCCL prints its version banner on fd 1: keep stdout for the JSON line only
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default="c2", choices=sorted(SCENES) + ["c5"])
    ap.add_argument("--triangles", type=int, default=10000000, help="c5: soup size")
    ap.add_argument("--soup-film", type=int, default=2048, help="c5: film resolution (square)")
    ap.add_argument("--passes", type=int, default=16, help="c5: jittered passes per step")
    ap.add_argument("--spp", type=int, default=None, help="iterations per step (default: the scene's sample count)")
    ap.add_argument("--partition", default="samples", choices=["samples", "tiles"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import __graft_entry__ as g
    if rank == 0 and not os.path.exists(os.path.join(ROOT, "pearray_b200", "libprb200.so")):
        g.build()
    if args.scene == "c5":
        if args.impl == "reference":
            emit({"impl": "reference", "unavailable": "c5 is a GPU ray-stream workload; the CPU arm is defined for the path-tracing configs c1-c4"})
            return
        run_soup(args, rank, world, local_rank)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
