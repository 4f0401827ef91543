#!/bin/bash
# one full ncu capture of the rollout kernel with per-instruction source counters (development aid / profiles/)
# usage under gpurun: bash scripts/gpu_ncu_rollout.sh <tag> [frames] [lanes]
cd "$(dirname "$0")/.."
tag=${1:-rollout}
python scripts/rollout_bench.py ${2:-256} ${3:-1024} 2 wstd=0.001 > gpurun_out/${tag}_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:trx_rollout_kernel -s 1 -c 1 -f -o gpurun_out/${tag} \
  python scripts/rollout_bench.py ${2:-256} ${3:-1024} 1 wstd=0.001 > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv 2>/dev/null
ls -la gpurun_out/${tag}*
cat gpurun_out/${tag}_plain.log | grep frames
