#!/bin/bash
# development aid: objective time vs. chain cap of the lowering (TRX_CHAIN)
for c in ${CAPS:-1 2 4 8 16 64}; do
  echo "== TRX_CHAIN=$c"
  TRX_CHAIN=$c BENCH_SKIP_CPU=1 timeout 300 python bench.py --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['loss'], {k:(v['n_levels'],v['n_bundles'],v.get('n_chained')) for k,v in d['lowering_stats'].items()})"
done
