#!/bin/sh
# round 2, N GPUs of one box (gpurun --gpus N): NCCL film reduce tests, bench.py under torchrun (weak: C2 by sample ranges; strong: complex.prc by tiles / sample ranges)
N=${NGPU:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02_gpu_tests_multi_n$N.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; cut -c1-300 gpurun_out/r02_bench_n$N.json; tail -3 gpurun_out/r02_bench_n$N.err
python - <<'P'
import json,os
N=os.environ.get('NGPU','8')
d=json.loads(open('gpurun_out/r02_bench_n%s.json'%N).read().strip().splitlines()[-1])
print('weak C2', d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['rank_ms_per_step'])
s=d.get('strong_scaling') or {}
print('strong tiles', s.get('value',0)/1e6, s.get('ms_per_step'), s.get('rank_ms_per_step'), s.get('tile_film_bit_identical_to_1gpu'))
b=s.get('by_sample_ranges') or {}
print('strong samples', json.dumps(b)[:600])
P
