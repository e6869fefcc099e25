#!/bin/sh
# round 2, call U: box-culled two-phase small-scene trace (traverseSmall) vs the exhaustive one (gpurun_variants/lib_head.so)
mkdir -p gpurun_out
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'])"; }
run() {
  python bench.py --scene c2 --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c0 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c1 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c3 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
}
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_u.log 2>&1; tail -4 gpurun_out/r02_gpu_tests_u.log
echo "== new (box-culled)"; run
cp pearray_b200/libprb200.so /tmp/lib_new.so
cp gpurun_variants/lib_head.so pearray_b200/libprb200.so
echo "== head (exhaustive)"; run
cp /tmp/lib_new.so pearray_b200/libprb200.so
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_trace" -s 30 -c 1 -o gpurun_out/r02_c2_u -f python bench.py --scene c2 --no-cpu --no-extras --steps 1 --warmup 1 --spp 16 > gpurun_out/ncu_c2_u.log 2>&1
ls -la gpurun_out/
