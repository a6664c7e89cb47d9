#!/bin/bash
# A/B of library variants (tools/build_variant.sh) on the search-latency probe: tools/ab_probe.sh "148 1184" old new ...
sizes=$1; shift
for v in "$@"; do
  for n in $sizes; do
    POASTA_B200_LIB=$PWD/variants/lib_$v.so POASTA_LANES=1 python bench.py --families $n --steps 1 --warmup 1 --seqs 8 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$v', $n, 'step_ms', round(d['ms_per_step']), 'search_ms', round(d['device_ms_per_step']['order_exact_search']), 'gcups', round(d['value'],3))"
  done
done
