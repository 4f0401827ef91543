#!/bin/bash
# BASELINE.json configs 2 and 5 (and 3 for reference) through the rollout objective on one GPU: one JSON line each
cd "$(dirname "$0")/.."
out=gpurun_out/${1:-r2}_configs.jsonl
rm -f $out
export ROLLOUT_JSON=$out ROLLOUT_MAX_GB=120
python scripts/rollout_bench.py 256 64 5 kind=hand8 hidden=0 2>&1 | grep "^frames"            # config 2
python scripts/rollout_bench.py 256 1024 5 kind=handarm hidden=64 2>&1 | grep "^frames"       # config 3
for r in 1024 4096 16384 65536; do                                                            # config 5 sweep
  python scripts/rollout_bench.py 1024 $r 2 kind=chain joints=30 contacts=16 hidden=16 2>&1 | grep "^frames"
done
python scripts/rollout_bench.py 256 8192 3 kind=handarm hidden=64 2>&1 | grep "^frames"       # configs[3] total on one GPU
