#!/bin/sh
# round 2, call N2: k_shade_geom over a compacted list of the live slots in every staged scene (k_compact_active before it), against the previous build
mkdir -p gpurun_out
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'], d.get('shading'))"; }
run() {
  python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | q
  python bench.py --scene c1 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c3 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  PRB_STAGED=1 python bench.py --scene c4c --no-cpu --steps 1 --warmup 1 --spp 32 2>/dev/null | q
}
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_n2.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_n2.log
echo "== new"; run
cp pearray_b200/libprb200.so /tmp/lib_new.so
cp gpurun_variants/lib_m2.so pearray_b200/libprb200.so
echo "== previous build"; run
cp /tmp/lib_new.so pearray_b200/libprb200.so
