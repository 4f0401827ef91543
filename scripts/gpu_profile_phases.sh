#!/bin/bash
# per-phase clock profile of the rollout kernel (development aid): rebuilds the library with
# -DTRX_ROLLOUT_PROFILE (plus $EXTRA) in the GPU box's scratch copy and runs scripts/rollout_bench.py
# usage under gpurun: [EXTRA="-D..."] bash scripts/gpu_profile_phases.sh [frames] [lanes] [key=value problem options ...]
set -e
cd "$(dirname "$0")/.."
TRX_NVCC_EXTRA="-DTRX_ROLLOUT_PROFILE $EXTRA" python -c "from robio_2021_dexopt_b200.build import build_native; build_native(force=True)" 2>&1 | grep -v "warning\|\^\|^$\|int32_t\|detected during" || true
python scripts/rollout_bench.py ${1:-256} ${2:-1024} 3 "${@:3}" 2>&1 | grep -v "Warn\|sqnorm"
