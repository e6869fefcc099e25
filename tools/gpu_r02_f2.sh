#!/bin/sh
# round 2, call F2: product of two leaf nodes evaluated in line (light radiance), against call E2 build
mkdir -p gpurun_out /tmp/reps
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'], d.get('shading'))"; }
run() {
  python bench.py --scene c2 --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c0 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c1 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c3 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
}
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_f2.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_f2.log
echo "== new"; run
cp pearray_b200/libprb200.so /tmp/lib_new.so
cp gpurun_variants/lib_e2.so pearray_b200/libprb200.so
echo "== call E2 build"; run
cp /tmp/lib_new.so pearray_b200/libprb200.so
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_shade" -s 30 -c 1 -o /tmp/reps/r02_c2_f2 -f python bench.py --scene c2 --no-cpu --no-extras --steps 1 --warmup 1 --spp 16 > gpurun_out/ncu_c2_f2.log 2>&1
python tools/ncu_summary.py /tmp/reps/r02_c2_f2.ncu-rep --all > gpurun_out/r02_ncu_c2_f2.txt 2>&1
