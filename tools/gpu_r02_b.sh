#!/bin/sh
# round 2, call B: full GPU suite, the default bench line with the new workload keys, k_shade occupancy variants, small-scene kernel A/B
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests_b.log 2>&1; tail -8 gpurun_out/r02_gpu_tests_b.log
timeout 900 python bench.py > gpurun_out/r02_bench_default_b.json 2> gpurun_out/r02_bench_default_b.err; cut -c1-300 gpurun_out/r02_bench_default_b.json; tail -3 gpurun_out/r02_bench_default_b.err
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'])"; }
echo "== small-scene kernel off (PRB_TRACE_MODE=bvh)"; PRB_TRACE_MODE=bvh python bench.py --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
echo "== base"; python bench.py --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
python bench.py --scene c3 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | q
cp pearray_b200/libprb200.so /tmp/lib_base.so
for v in minb5 minb6 minb8; do
  cp gpurun_variants/lib_$v.so pearray_b200/libprb200.so
  echo "== variant $v"
  python bench.py --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c3 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | q
done
cp /tmp/lib_base.so pearray_b200/libprb200.so
