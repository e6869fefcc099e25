#!/bin/sh
# round 2, call I2: host-buffer ray streams in chunks on two CUDA streams (upload / traverse / download overlapped): C5 e2e; hit parity tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "hits or soup or smoke or camera or invalid" > gpurun_out/r02_gpu_tests_i2.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_i2.log
python bench.py --scene c5 --steps 1 --warmup 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5', round(d['value']/1e6,1), d['e2e'], d.get('parity_vs_oracle'), d.get('resident_equals_host_path'))"
cp pearray_b200/libprb200.so /tmp/lib_new.so
cp gpurun_variants/lib_h2.so pearray_b200/libprb200.so
python bench.py --scene c5 --no-cpu --steps 1 --warmup 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5 (call H2 build)', round(d['value']/1e6,1), d['e2e'])"
cp /tmp/lib_new.so pearray_b200/libprb200.so
python -c "import __graft_entry__ as g; g.smoke()"
