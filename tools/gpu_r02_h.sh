#!/bin/sh
# round 2, call H: are the per-warp statistics atomics (7 x 64-bit RED per warp into one 128-byte line) the floor of k_shade?
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'])"; }
run() {
  python bench.py --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c3 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | q
  python bench.py --scene c4c --no-cpu --steps 1 --warmup 1 --spp 16 2>/dev/null | q
}
echo "== base"; run
cp pearray_b200/libprb200.so /tmp/lib_base.so
cp gpurun_variants/lib_nostats.so pearray_b200/libprb200.so
echo "== statistics atomics compiled out"; run
cp /tmp/lib_base.so pearray_b200/libprb200.so
