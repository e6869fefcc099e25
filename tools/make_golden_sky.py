#!/usr/bin/env python3
"""Golden vectors for the Hosek-Wilkie sky model, produced by the REFERENCE's own C sources
(/root/reference/src/skysun/skysun/model/ArHosekSkyModel.cpp compiled as-is into oracle/_ref/libarhosek.so by
oracle/Makefile).  Run in the build container; the output tests/golden/hosek_reference.npz travels to the GPU box.

  inputs   (n, 6) float64: solar elevation [rad], turbidity, ground albedo, theta, gamma, wavelength [nm]
  radiance (n,)   float64: arhosekskymodel_radiance(arhosekskymodelstate_alloc_init(se, turbidity, albedo), theta, gamma, wl)
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "oracle", "_ref", "libarhosek.so")
subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_ref/libarhosek.so"])
lib = C.CDLL(so)
syms = [l.split()[-1] for l in subprocess.run(["nm", "-D", so], capture_output=True, text=True).stdout.splitlines()]
init = getattr(lib, [s for s in syms if "arhosekskymodelstate_alloc_init" in s and "alien" not in s][0])
rad = getattr(lib, [s for s in syms if "arhosekskymodel_radiance" in s][0])
free = getattr(lib, [s for s in syms if "arhosekskymodelstate_free" in s][0])
init.restype, init.argtypes = C.c_void_p, [C.c_double] * 3
rad.restype, rad.argtypes = C.c_double, [C.c_void_p] + [C.c_double] * 3
free.argtypes = [C.c_void_p]

rng = np.random.default_rng(20261017)
n = 4096
x = np.stack([rng.uniform(0.0, np.pi / 2, n), rng.uniform(1.0, 10.0, n), rng.uniform(0.0, 1.0, n), rng.uniform(0.0, np.pi / 2, n),
              rng.uniform(0.0, np.pi, n), rng.uniform(320.0, 760.0, n)], axis=1)
x[:64, 1] = np.repeat(np.arange(1, 9), 8)  # integer turbidities (and 10 below) take the corner branches
x[64:96, 1] = 10.0
x[96:128, 5] = 320.0 + 40.0 * rng.integers(0, 11, 32)  # wavelengths on band centres: no interpolation
y = np.empty(n)
for i, (se, tb, al, th, ga, wl) in enumerate(x):
    st = init(se, tb, al)
    y[i] = rad(st, th, ga, wl)
    free(st)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hosek_reference.npz"), inputs=x, radiance=y)
print("hosek_reference: n", n, "mean", y.mean(), "max", y.max())
