#!/bin/sh
# round 2, call C2: list atomics of k_shade_geom issued at once (match_any), against call B2's build (gpurun_variants/lib_b2.so)
mkdir -p gpurun_out /tmp/reps
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'], d.get('shading'))"; }
run() {
  python bench.py --scene c3 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | q
  python bench.py --scene c4c --no-cpu --steps 1 --warmup 1 --spp 16 2>/dev/null | q
}
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_c2.log 2>&1; tail -2 gpurun_out/r02_gpu_tests_c2.log
echo "== new"; run
cp pearray_b200/libprb200.so /tmp/lib_new.so
cp gpurun_variants/lib_b2.so pearray_b200/libprb200.so
echo "== call B2 build"; run
cp /tmp/lib_new.so pearray_b200/libprb200.so
