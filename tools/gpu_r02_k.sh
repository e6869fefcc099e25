#!/bin/sh
# round 2, call K: shading path chosen by measurement (auto) vs pinned single / staged; full GPU test suite
mkdir -p gpurun_out
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'], d.get('shading'))"; }
run() {
  python bench.py --scene c3 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | q
  python bench.py --scene c4c --no-cpu --steps 1 --warmup 1 --spp 16 2>/dev/null | q
  python bench.py --scene c1 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c0 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
}
echo "== auto"; run
echo "== PRB_STAGED=0"; export PRB_STAGED=0; run
echo "== PRB_STAGED=1"; export PRB_STAGED=1; run
unset PRB_STAGED
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_k.log 2>&1; tail -4 gpurun_out/r02_gpu_tests_k.log
