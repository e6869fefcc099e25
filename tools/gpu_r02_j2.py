"""call J2: where does the time of a host-buffer ray stream go?  link bandwidth (pinned H2D / D2H), prb_trace_closest with page-locked
columns for several stream lengths, kernel-only time of the same rays"""
import time, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import pearray_b200 as prb
dev = torch.device("cuda:0")
for mb in (1, 4, 32):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory(); d = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    for name, a, b in (("H2D", h, d), ("D2H", d, h)):
        b.copy_(a, non_blocking=True); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20): b.copy_(a, non_blocking=True)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
        print("%s %2d MB: %.1f GB/s (%.0f us)" % (name, mb, (mb << 20) / dt / 1e9, dt * 1e6))
scene = prb.Scene.from_file("scenes/c4_boltsandgears.prc")
ctx = prb.Context(0); ctx.upload_scene(scene); ctx.upload_rng(scene.rng_map())
org, dr, wvl, pix = ctx.generate_camera_rays([(0, 0, scene.width, scene.height)], 0)
n = len(org); print("rays", n)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
cols = [pin(org[:, i]) for i in range(3)] + [pin(dr[:, i]) for i in range(3)] + [None, None]
out = [pin(np.empty(n, np.uint32)), pin(np.empty(n, np.uint32)), pin(np.empty(n, np.float32)), pin(np.empty(n, np.float32)), pin(np.empty(n, np.float32))]
for m in (1 << 18, 1 << 19, n):
    c = [None if x is None else x[:m] for x in cols]; o = [x[:m] for x in out]
    ctx.trace_closest_soa(c, m, o)
    t0 = time.perf_counter()
    for _ in range(10): ctx.trace_closest_soa(c, m, o)
    dt = (time.perf_counter() - t0) / 10
    print("host columns n=%7d: %.3f ms wall, %.3f ms on the stream, %.0f Mrays/s, %.1f GB/s of copies" % (m, dt * 1e3, ctx.last_device_ms(), m / dt / 1e6, m * 44 / dt / 1e9))
