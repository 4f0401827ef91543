#!/bin/bash
# A/B timing of prebuilt library variants (development aid): _variants/lib*.so are loaded through TRX_B200_LIB
# usage under gpurun: bash scripts/gpu_ab_variants.sh <tag> <variant to test on the GPU suite, or -> <variants to time ...>
cd "$(dirname "$0")/.."
tag=$1; check=$2; shift 2
if [ "$check" != "-" ]; then
  TRX_B200_LIB=$PWD/_variants/lib$check.so timeout 240 python -m pytest tests/test_rollout.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${tag}_test_$check.log
  cat gpurun_out/${tag}_test_$check.log
fi
for v in "$@"; do
  for rep in 1 2; do
    echo "variant $v" >> gpurun_out/${tag}_ab.txt
    TRX_B200_LIB=$PWD/_variants/lib$v.so timeout 120 python scripts/rollout_bench.py 256 1024 20 wstd=0.001 2>&1 | grep "per evaluation" >> gpurun_out/${tag}_ab.txt
  done
done
cat gpurun_out/${tag}_ab.txt
