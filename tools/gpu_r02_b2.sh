#!/bin/sh
# round 2, call B2: geometry point in line in k_shade_geom and the generic k_shade, against call A2's build (gpurun_variants/lib_a2.so)
mkdir -p gpurun_out /tmp/reps
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'], d.get('shading'))"; }
run() {
  python bench.py --scene c3 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | q
  python bench.py --scene c4c --no-cpu --steps 1 --warmup 1 --spp 16 2>/dev/null | q
  python bench.py --scene c1 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
}
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_b2.log 2>&1; tail -2 gpurun_out/r02_gpu_tests_b2.log
echo "== new"; run
cp pearray_b200/libprb200.so /tmp/lib_new.so
cp gpurun_variants/lib_a2.so pearray_b200/libprb200.so
echo "== call A2 build"; run
cp /tmp/lib_new.so pearray_b200/libprb200.so
PRB_STAGED=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_shade|k_trace" -s 60 -c 9 -o /tmp/reps/r02_c4 -f python bench.py --scene c4 --no-cpu --no-extras --steps 1 --warmup 1 --spp 8 > gpurun_out/ncu_c4.log 2>&1
python tools/ncu_summary.py /tmp/reps/r02_c4.ncu-rep --all > gpurun_out/r02_ncu_c4_b2.txt 2>&1
python tools/ncu_hotspots.py /tmp/reps/r02_c4.ncu-rep k_shade_geom pearray_b200/libprb200.so 60 k_shade_geomILb0E > gpurun_out/r02_hotspots_c4_geom_b2.txt 2>&1
ls -la /tmp/reps
