#!/bin/sh
# round 2, call T: staged kernels templated on LPE; full GPU tests
mkdir -p gpurun_out
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'])"; }
python bench.py --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | q
python bench.py --scene c3 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
python bench.py --scene c4c --no-cpu --steps 1 --warmup 1 --spp 16 2>/dev/null | q
python bench.py --scene c1 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_t.log 2>&1; tail -4 gpurun_out/r02_gpu_tests_t.log
