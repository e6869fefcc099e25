#!/bin/sh
# round 2, call O: what k_shade<128,1,Lambert> lost between fd5c646 (270 M) and the build with LPE / textures / AOV ext (253 M)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_shade|k_trace" -s 60 -c 2 -o gpurun_out/r02_c2_q -f python bench.py --scene c2 --no-cpu --no-extras --steps 1 --warmup 1 --spp 16 > gpurun_out/ncu_c2.log 2>&1
tail -3 gpurun_out/ncu_c2.log
