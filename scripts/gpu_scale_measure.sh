#!/bin/bash
# multi-GPU bench lines (run under gpurun --gpus 8): weak scaling (1024 rollouts per GPU, BASELINE configs[3] at 8)
# and strong scaling (8192 rollouts in total) -- one JSON line per run under gpurun_out/
cd "$(dirname "$0")/.."
tag=${1:-r2}
run() {  # n_gpus, name, env...
  n=$1; name=$2; shift 2
  if [ "$n" = 1 ]; then
    env BENCH_SKIP_CPU=1 "$@" python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  else
    env BENCH_SKIP_CPU=1 "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29$((600 + n)) \
      bench.py --gpus $n --steps 10 --warmup 3 2> gpurun_out/${tag}_${name}.err | grep '^{' > gpurun_out/${tag}_${name}.json
  fi
  python -c "import json,sys; d=json.load(open('gpurun_out/${tag}_${name}.json')); print('${name}', d['n_gpus'], '%.2f M/s' % (d['value']/1e6), '%.3f ms' % d['ms_per_step'], 'e2e %.2f M/s' % (d['e2e']['value']/1e6))"
}
run 4 weak_n4
run 8 weak_n8
run 1 strong_n1 BENCH_TOTAL_ROLLOUTS=8192
run 2 strong_n2 BENCH_TOTAL_ROLLOUTS=8192
run 4 strong_n4 BENCH_TOTAL_ROLLOUTS=8192
run 8 strong_n8 BENCH_TOTAL_ROLLOUTS=8192
