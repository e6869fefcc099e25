#!/bin/sh
# round 2, call T2: blocks of k_trace_small without work leave before they copy the scene, against the previous build
mkdir -p gpurun_out
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'])"; }
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_t2.log 2>&1; tail -2 gpurun_out/r02_gpu_tests_t2.log
echo "== new"; python bench.py --scene c2 --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q; python bench.py --scene c0 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
cp pearray_b200/libprb200.so /tmp/lib_new.so; cp gpurun_variants/lib_final.so pearray_b200/libprb200.so
echo "== previous"; python bench.py --scene c2 --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q; python bench.py --scene c0 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
cp /tmp/lib_new.so pearray_b200/libprb200.so
