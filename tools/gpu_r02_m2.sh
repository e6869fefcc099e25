#!/bin/sh
# round 2, call M2: what does FMA contraction (-fmad=true) buy in time and cost in the film?  (structural limits)
# The shipped build is -fmad=false: its films are bit-identical to the oracle.  Same seeds, same scenes, both builds.
mkdir -p gpurun_out
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'])"; }
run() {
  python bench.py --scene c2 --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c3 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | q
  python bench.py --scene c4c --no-cpu --steps 1 --warmup 1 --spp 16 2>/dev/null | q
}
echo "== -fmad=false (shipped)"; run
for s in c2_cornellbox c3_cornellbox_glassy c4_boltsandgears; do python tools/film_dump.py scenes/$s.prc 64 /tmp/${s}_ieee.npy; done
python tools/film_dump.py scenes/c2_cornellbox.prc 1024 /tmp/c2_1024_ieee.npy
cp pearray_b200/libprb200.so /tmp/lib_new.so
cp gpurun_variants/lib_fma.so pearray_b200/libprb200.so
echo "== -fmad=true"; run
for s in c2_cornellbox c3_cornellbox_glassy c4_boltsandgears; do python tools/film_dump.py scenes/$s.prc 64 /tmp/${s}_fma.npy; done
python tools/film_dump.py scenes/c2_cornellbox.prc 1024 /tmp/c2_1024_fma.npy
cp /tmp/lib_new.so pearray_b200/libprb200.so
python - <<'P'
import numpy as np
def rel(a, b): return float(np.sqrt(np.mean((a.astype(np.float64) - b) ** 2)) / np.mean(np.abs(b)))
for s in ("c2_cornellbox", "c3_cornellbox_glassy", "c4_boltsandgears", "c2_1024"):
    a, b = np.load("/tmp/%s_fma.npy" % s), np.load("/tmp/%s_ieee.npy" % s)
    ra, rb = np.load("/tmp/%s_fma_rng.npy" % s), np.load("/tmp/%s_ieee_rng.npy" % s)
    print("%-22s relRMSE(fma film vs shipped film) %.3e   bit-identical pixels %.1f %%   pixels with the same RNG state after the render %.1f %%"
          % (s, rel(a, b), 100.0 * np.mean(np.all(a.view(np.uint32) == b.view(np.uint32), axis=-1)), 100.0 * np.mean(ra == rb)))
P
