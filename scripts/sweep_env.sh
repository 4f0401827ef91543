#!/bin/bash
# development aid: bench (no CPU baseline) under a list of environment settings, one per line of $1
while IFS= read -r line; do
  [ -z "$line" ] && continue
  echo "== $line"
  env $line BENCH_SKIP_CPU=1 timeout 300 python bench.py --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['loss'], {k:(v['n_levels'],v['n_bundles'],v.get('n_chained'),v.get('n_folded')) for k,v in d['lowering_stats'].items()})"
done < "$1"
