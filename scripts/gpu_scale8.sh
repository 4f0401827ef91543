#!/bin/bash
# 8-GPU weak-scaling bench line (1024 rollouts per GPU) -- run under gpurun --gpus 8
cd "$(dirname "$0")/.."
env BENCH_SKIP_CPU=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 bench.py --gpus 8 --steps 10 --warmup 3 2> gpurun_out/r2j_weak_n8.err | grep '^{' > gpurun_out/r2j_weak_n8.json
python -c "import json; d=json.load(open('gpurun_out/r2j_weak_n8.json')); print('weak_n8', d['n_gpus'], '%.2f M/s' % (d['value']/1e6), '%.3f ms' % d['ms_per_step'], 'e2e %.2f M/s' % (d['e2e']['value']/1e6))"
