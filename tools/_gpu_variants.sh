mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/gpu_tests.log 2>&1; tail -8 gpurun_out/gpu_tests.log
for sc in "c1 16" "c2 256" "c3 128" "c4 64" "c4c 16"; do set -- $sc
  timeout 300 python bench.py --scene $1 --no-cpu --steps 1 --warmup 1 --spp $2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,1), d['stage_ms'])"
done
