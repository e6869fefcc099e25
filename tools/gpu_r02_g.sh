#!/bin/sh
# round 2, call G: node-value cache A/B; k_shade with a smaller executed footprint (NEE off) under ncu
mkdir -p gpurun_out
q() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['stage_ms'])"; }
run() {
  python bench.py --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c3 --no-cpu --steps 2 --warmup 1 2>/dev/null | q
  python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | q
}
echo "== node cache on (HEAD)"; run
cp pearray_b200/libprb200.so /tmp/lib_base.so
cp gpurun_variants/lib_nocache.so pearray_b200/libprb200.so
echo "== node cache off"; run
cp /tmp/lib_base.so pearray_b200/libprb200.so
cp scenes/c2_cornellbox.prc /tmp/c2.prc; cp gpurun_variants/c2_nonee.prc scenes/c2_cornellbox.prc
echo "== C2 with :nee false"; python bench.py --no-cpu --no-extras --steps 2 --warmup 1 2>/dev/null | q
timeout 300 ncu --set full --clock-control none -k regex:"k_shade" -s 30 -c 1 -o gpurun_out/r02_c2_nonee -f python bench.py --scene c2 --no-cpu --no-extras --steps 1 --warmup 1 --spp 16 > gpurun_out/ncu_c2.log 2>&1
cp /tmp/c2.prc scenes/c2_cornellbox.prc
timeout 300 ncu --set full --clock-control none -k regex:"k_shade" -s 30 -c 1 -o gpurun_out/r02_c2_head -f python bench.py --scene c2 --no-cpu --no-extras --steps 1 --warmup 1 --spp 16 > gpurun_out/ncu_c2.log 2>&1
