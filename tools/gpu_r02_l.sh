#!/bin/sh
# round 2, call L: LPE channels, extended AOVs, film reduce of the extras: full GPU test suite + default bench smoke
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_l.log 2>&1; tail -15 gpurun_out/r02_gpu_tests_l.log
python bench.py --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'])"
