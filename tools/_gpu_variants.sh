mkdir -p gpurun_out
cp pearray_b200/libprb200.so /tmp/lib_base.so
for v in base m5 m6 m8; do
  if [ $v != base ]; then cp gpurun_variants/lib_$v.so pearray_b200/libprb200.so; else cp /tmp/lib_base.so pearray_b200/libprb200.so; fi
  echo "== variant $v"
  timeout 200 python bench.py --scene c2 --no-cpu --steps 1 --warmup 1 --spp 256 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2', round(d['value']/1e6,1), d['stage_ms'])"
  timeout 200 python bench.py --scene c4c --no-cpu --steps 1 --warmup 1 --spp 16 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4c', round(d['value']/1e6,1), d['stage_ms'])"
done
cp /tmp/lib_base.so pearray_b200/libprb200.so
