#!/bin/sh
# round 2, call E (2 GPUs): NCCL film reduce of the C ABI, bench.py under torchrun
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_variance_aov.py tests/test_plugin_abi.py -m gpu -q > gpurun_out/r02_gpu_tests_multi_e.log 2>&1; tail -6 gpurun_out/r02_gpu_tests_multi_e.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_n2_e.json 2> gpurun_out/r02_bench_n2_e.err; cut -c1-400 gpurun_out/r02_bench_n2_e.json; tail -5 gpurun_out/r02_bench_n2_e.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 2 --warmup 1 --partition tiles --no-extras > gpurun_out/r02_bench_n2_tiles_e.json 2> gpurun_out/r02_bench_n2_tiles_e.err; cut -c1-300 gpurun_out/r02_bench_n2_tiles_e.json
