#!/bin/sh
# round 2, call F: fused trace || shade schedule A/B (PRB_FUSED=0/1), node-value cache, bit-exactness of both schedules
mkdir -p gpurun_out
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'], d['roofline']['kernel'], round(d['roofline']['frac'],4))"; }
for f in 0 1; do
  echo "== PRB_FUSED=$f"
  PRB_FUSED=$f python bench.py --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
  PRB_FUSED=$f python bench.py --scene c1 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  PRB_FUSED=$f python bench.py --scene c0 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
done
PRB_FUSED=1 timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_f_fused.log 2>&1; tail -4 gpurun_out/r02_gpu_tests_f_fused.log
PRB_FUSED=0 timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_f.log 2>&1; tail -4 gpurun_out/r02_gpu_tests_f.log
PRB_FUSED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_fused" -s 20 -c 2 -o gpurun_out/r02_c2_fused -f python bench.py --scene c2 --no-cpu --no-extras --steps 1 --warmup 1 --spp 16 > gpurun_out/ncu_c2.log 2>&1
