#!/bin/sh
# round 2, call I: staged shading (k_shade_geom / _nee / _scatter) vs the single k_shade; bit-exactness under both
mkdir -p gpurun_out
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'])"; }
run() {
  python bench.py --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c3 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | q
  python bench.py --scene c4c --no-cpu --steps 1 --warmup 1 --spp 16 2>/dev/null | q
  python bench.py --scene c1 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
}
echo "== PRB_STAGED=0"; export PRB_STAGED=0; run
echo "== PRB_STAGED=1"; export PRB_STAGED=1; run
PRB_STAGED=1 timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_i_staged.log 2>&1; tail -4 gpurun_out/r02_gpu_tests_i_staged.log
PRB_STAGED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_shade" -s 30 -c 3 -o gpurun_out/r02_c2_staged -f python bench.py --scene c2 --no-cpu --no-extras --steps 1 --warmup 1 --spp 16 > gpurun_out/ncu_c2.log 2>&1
PRB_STAGED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_shade" -s 30 -c 3 -o gpurun_out/r02_c4_staged -f python bench.py --scene c4 --no-cpu --no-extras --steps 1 --warmup 1 --spp 8 > gpurun_out/ncu_c4.log 2>&1
