#!/bin/sh
# round 2, call A2: leaf nodes evaluated in line in the Lambert kernels (evalNodeFast), against call Z build
mkdir -p gpurun_out /tmp/reps
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'], d.get('shading'))"; }
run() {
  python bench.py --scene c2 --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c0 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c1 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c3 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | q
}
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_a2.log 2>&1; tail -2 gpurun_out/r02_gpu_tests_a2.log
echo "== new"; run
cp pearray_b200/libprb200.so /tmp/lib_new.so
cp gpurun_variants/lib_z.so pearray_b200/libprb200.so
echo "== call Z build"; run
cp /tmp/lib_new.so pearray_b200/libprb200.so
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_shade" -s 30 -c 1 -o /tmp/reps/r02_c2_a2 -f python bench.py --scene c2 --no-cpu --no-extras --steps 1 --warmup 1 --spp 16 > gpurun_out/ncu_c2_a2.log 2>&1
python tools/ncu_summary.py /tmp/reps/r02_c2_a2.ncu-rep --all > gpurun_out/r02_ncu_c2_a2.txt 2>&1
python tools/ncu_hotspots.py /tmp/reps/r02_c2_a2.ncu-rep k_shade pearray_b200/libprb200.so 400 k_shadeILi128ELi1ELi2E > gpurun_out/r02_hotspots_c2_shade_a2.txt 2>&1
cp /tmp/reps/r02_c2_a2.ncu-rep gpurun_out/
