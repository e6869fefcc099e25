#!/bin/sh
# round 2, call A: GPU test-suite, baseline bench lines, fresh ncu captures (launch list + --set full of the hot kernels)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > gpurun_out/r02_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_a.log 2>&1; tail -5 gpurun_out/r02_gpu_tests_a.log
timeout 300 python bench.py --no-cpu > gpurun_out/r02_bench_c2_a.json 2> gpurun_out/r02_bench_c2_a.err; cut -c1-400 gpurun_out/r02_bench_c2_a.json
timeout 300 python bench.py --scene c5 --no-cpu --steps 1 > gpurun_out/r02_bench_c5_a.json 2>/dev/null; cut -c1-300 gpurun_out/r02_bench_c5_a.json
timeout 300 python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 > gpurun_out/r02_bench_c4_a.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_c4_a.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 3 -o gpurun_out/r02_c5_a -f python bench.py --scene c5 --no-cpu --steps 1 --warmup 1 --passes 1 > gpurun_out/ncu_c5.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_shade|k_trace|k_regen" -s 90 -c 3 -o gpurun_out/r02_c2_a -f python bench.py --scene c2 --no-cpu --steps 1 --warmup 1 --spp 16 > gpurun_out/ncu_c2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_shade|k_trace|k_regen" -s 90 -c 3 -o gpurun_out/r02_c4_a -f python bench.py --scene c4 --no-cpu --steps 1 --warmup 1 --spp 8 > gpurun_out/ncu_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
