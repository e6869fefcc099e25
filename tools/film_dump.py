"""Renders `spp` iterations of a scene with the library that is currently in place and saves the film (XYZ, filtered) as .npy --
used to compare the films of two builds (tools/gpu_r02_m2.sh: the -fmad=true experiment)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pearray_b200 as prb
scene_file, spp, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
scene = prb.Scene.from_file(scene_file)
ctx = prb.Context(0)
ctx.upload_scene(scene)
ctx.upload_rng(scene.rng_map())
ctx.render_tiles([(0, 0, scene.width, scene.height)], 0, spp)
xyz, cnt = ctx.film()
np.save(out, xyz)
np.save(out.replace(".npy", "_rng.npy"), ctx.download_rng())
