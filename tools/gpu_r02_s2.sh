#!/bin/sh
# round 2, call S2: all-Lambert small scenes walk a compacted list of the slots that still have work (k_compact_small); PRB_COMPACT_SMALL=0 is the previous behaviour
mkdir -p gpurun_out
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'])"; }
run() {
  python bench.py --scene c2 --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c0 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
}
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_s2.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_s2.log
echo "== compacted"; run
export PRB_COMPACT_SMALL=0
echo "== PRB_COMPACT_SMALL=0"; run
