mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/gpu_tests.log 2>&1; tail -8 gpurun_out/gpu_tests.log
cp pearray_b200/libprb200.so /tmp/lib_b16.so
for v in 16 1; do
  if [ $v != 16 ]; then cp gpurun_variants/lib_b$v.so pearray_b200/libprb200.so; else cp /tmp/lib_b16.so pearray_b200/libprb200.so; fi
  echo "== variant $v"
  timeout 200 python bench.py --scene c5 --no-cpu --steps 1 --warmup 1 --passes 8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5', {k:round(v['mrays_per_s']) for k,v in d['classes'].items()})"
  timeout 200 python bench.py --scene c2 --no-cpu --steps 1 --warmup 1 --spp 256 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2', round(d['value']/1e6,1), d['stage_ms'])"
  timeout 200 python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 --spp 64 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4', round(d['value']/1e6,1), d['stage_ms'])"
done
cp /tmp/lib_b16.so pearray_b200/libprb200.so
