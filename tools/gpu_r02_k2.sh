#!/bin/sh
# round 2, call K2: C5 host-buffer leg on a coherent ray block: chunked two-stream prb_trace_* against call H2's build (one stream)
python bench.py --scene c5 --steps 1 --warmup 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5', round(d['value']/1e6,1), d['e2e'], d.get('parity_vs_oracle'), d.get('resident_equals_host_path'), d['cpu_baseline'])"
cp pearray_b200/libprb200.so /tmp/lib_new.so
cp gpurun_variants/lib_h2.so pearray_b200/libprb200.so
python bench.py --scene c5 --no-cpu --steps 1 --warmup 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5 (call H2 build)', round(d['value']/1e6,1), d['e2e'])"
cp /tmp/lib_new.so pearray_b200/libprb200.so
