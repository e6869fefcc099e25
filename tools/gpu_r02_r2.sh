#!/bin/sh
# round 2, call R2: path rays of the slots that were active already run before the regeneration barrier of k_trace_small, against call P2's build
mkdir -p gpurun_out
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'])"; }
run() {
  python bench.py --scene c2 --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c0 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c3 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c1 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
}
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_r2.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_r2.log
echo "== new"; run
cp pearray_b200/libprb200.so /tmp/lib_new.so
cp gpurun_variants/lib_p2.so pearray_b200/libprb200.so
echo "== previous build"; run
cp /tmp/lib_new.so pearray_b200/libprb200.so
