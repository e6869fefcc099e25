#!/bin/sh
python tools/gpu_r02_j2.py
cp pearray_b200/libprb200.so /tmp/lib_new.so; cp gpurun_variants/lib_h2.so pearray_b200/libprb200.so
echo "== call H2 build (one stream, no chunks)"; python tools/gpu_r02_j2.py | grep "host columns"
cp /tmp/lib_new.so pearray_b200/libprb200.so
