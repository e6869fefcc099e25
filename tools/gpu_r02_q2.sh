#!/bin/sh
# round 2, call Q2: block size of k_trace_small (64 / 128 shipped / 256 threads)
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'])"; }
run() {
  python bench.py --scene c2 --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c0 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
}
cp pearray_b200/libprb200.so /tmp/lib_keep.so
for v in p2 ts64 ts256; do cp gpurun_variants/lib_$v.so pearray_b200/libprb200.so; echo "== $v"; run; done
cp /tmp/lib_keep.so pearray_b200/libprb200.so
