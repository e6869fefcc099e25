#!/bin/sh
# round 2, call O2: one hit bin in the all-Lambert kernels (no entity / material lookups for the sort key); what the statistics counters cost
mkdir -p gpurun_out
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'])"; }
run() {
  python bench.py --scene c2 --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c0 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
}
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_o2.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_o2.log
echo "== new (uniform sort key)"; run
cp pearray_b200/libprb200.so /tmp/lib_new.so
cp gpurun_variants/lib_final1.so pearray_b200/libprb200.so
echo "== previous build"; run
cp gpurun_variants/lib_nostats.so pearray_b200/libprb200.so
echo "== new, statistics counters compiled out (experiment)"; run
cp /tmp/lib_new.so pearray_b200/libprb200.so
