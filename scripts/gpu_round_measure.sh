#!/bin/bash
# the round's single-GPU measurement set (run under gpurun): GPU tests, the bench line with its CPU baseline, the
# reference arm, the ncu launch list of the bench command and one full capture of the rollout kernel
cd "$(dirname "$0")/.."
tag=${1:-r2}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${tag}_gputest.log
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench_n1.err
BENCH_SKIP_CPU=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 > gpurun_out/${tag}_launches_bench.log 2>&1
bash scripts/gpu_ncu_rollout.sh ${tag}_full > gpurun_out/${tag}_full.log 2>&1
cat gpurun_out/${tag}_gputest.log
head -c 400 gpurun_out/${tag}_bench_n1.json; echo
head -c 300 gpurun_out/${tag}_bench_reference.json; echo
tail -2 gpurun_out/${tag}_full.log
