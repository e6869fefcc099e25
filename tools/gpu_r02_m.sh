#!/bin/sh
# round 2, call M: image textures + environment map, LPE, extended AOVs: GPU test suite; C2 / C4 after the code growth
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_m.log 2>&1; tail -15 gpurun_out/r02_gpu_tests_m.log
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'])"; }
python bench.py --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | q
python bench.py --scene c3 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
