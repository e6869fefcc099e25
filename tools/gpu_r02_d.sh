#!/bin/sh
# round 2, call D: phase-structured k_shade, block size / phase barrier variants (instruction-cache sharing experiment)
mkdir -p gpurun_out
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'])"; }
cp pearray_b200/libprb200.so /tmp/lib_base.so
run() {
  python bench.py --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c3 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | q
  python bench.py --scene c4c --no-cpu --steps 1 --warmup 1 --spp 16 2>/dev/null | q
}
echo "== base (phases, no sync, 128)"; run
for v in blk512 sync128 sync256 sync512; do
  cp gpurun_variants/lib_$v.so pearray_b200/libprb200.so
  echo "== variant $v"; run
done
cp gpurun_variants/lib_sync512.so pearray_b200/libprb200.so
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_shade" -s 30 -c 1 -o gpurun_out/r02_c2_sync512 -f python bench.py --scene c2 --no-cpu --no-extras --steps 1 --warmup 1 --spp 16 > gpurun_out/ncu_c2.log 2>&1
cp /tmp/lib_base.so pearray_b200/libprb200.so
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_d.log 2>&1; tail -5 gpurun_out/r02_gpu_tests_d.log
