mkdir -p gpurun_out
cp pearray_b200/libprb200.so /tmp/lib_base.so
for v in base u256; do
  if [ $v != base ]; then cp gpurun_variants/lib_$v.so pearray_b200/libprb200.so; else cp /tmp/lib_base.so pearray_b200/libprb200.so; fi
  echo "== variant $v"
  for i in 1 2; do timeout 200 python bench.py --scene c2 --no-cpu --steps 1 --warmup 1 --spp 256 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2', round(d['value']/1e6,1), d['stage_ms'])"; done
done
cp /tmp/lib_base.so pearray_b200/libprb200.so
