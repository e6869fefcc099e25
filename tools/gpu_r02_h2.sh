#!/bin/sh
# round 2, call H2: k_shade_geom over the compacted active list (persistent trace): complex.prc single vs staged vs auto; new many-faces small-scene test
mkdir -p gpurun_out
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'], d.get('shading'))"; }
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_h2.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_h2.log
for m in auto 0 1; do
  if [ $m = auto ]; then unset PRB_STAGED; else export PRB_STAGED=$m; fi
  echo "== PRB_STAGED=$m"
  python bench.py --scene c4c --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | q
done
unset PRB_STAGED
python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | q
