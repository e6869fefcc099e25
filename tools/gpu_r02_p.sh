#!/bin/sh
# round 2, call R: FMA-form intersection test (Embree's msub / madd), leaf-only node evaluation and light sampling in the
# all-Lambert kernels; tests + C2 / C4 / C3 / complex / C5
mkdir -p gpurun_out
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'])"; }
python bench.py --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | q
python bench.py --scene c3 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
python bench.py --scene c4c --no-cpu --steps 1 --warmup 1 --spp 16 2>/dev/null | q
python bench.py --scene c5 --no-cpu --steps 2 --warmup 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k: round(v['mrays_per_s']) for k,v in d['classes'].items()})"
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_r.log 2>&1; tail -8 gpurun_out/r02_gpu_tests_p.log
