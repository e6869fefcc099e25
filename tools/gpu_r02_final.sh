#!/bin/sh
# round 2, final check of the shipped build the way the driver runs it: GPU test suite, smoke(), default bench line
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_final.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py > gpurun_out/r02_bench_default_final.json 2> gpurun_out/r02_bench_default_final.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_default_final.json').read().strip().splitlines()[-1])
print('C2 %.1f M  e2e %.1f M  ms/step %.1f  launches %d  clocks %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], d['gpu_launches'], d['clocks']))
print('roofline', {k: d['roofline'][k] for k in ('kernel','achieved','frac','traffic','avg_launch_us')}, 'cpu', d['cpu_baseline']['value'])
for k, v in d['workloads'].items(): print(k, round(v['value']/1e6,1), 'e2e', round(v['e2e']['value']/1e6,1), 'frac', round(v['roofline']['frac'],3), 'cpu', (v.get('cpu_baseline') or {}).get('value'), v.get('parity_vs_oracle'))
P
