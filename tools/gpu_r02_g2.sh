#!/bin/sh
# round 2, call G2: active slots compacted for the persistent k_trace (k_compact_active), against call F2's build (gpurun_variants/lib_f2.so)
mkdir -p gpurun_out /tmp/reps
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'], d.get('shading'))"; }
run() {
  python bench.py --scene c4c --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | q
  PRB_TRACE_MODE=persistent python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | q
}
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_g2.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_g2.log
echo "== new"; run
cp pearray_b200/libprb200.so /tmp/lib_new.so
cp gpurun_variants/lib_f2.so pearray_b200/libprb200.so
echo "== call F2 build"; run
cp /tmp/lib_new.so pearray_b200/libprb200.so
