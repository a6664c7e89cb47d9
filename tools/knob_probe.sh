#!/bin/bash
# Full-size exact-mode step under different runtime knobs: tools/knob_probe.sh "VAR=val VAR2=val" "..." ...
for cfg in "$@"; do
  env $cfg python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$cfg', 'step_ms', round(d['ms_per_step']), 'search_ms_sum', round(d['device_ms_per_step']['order_exact_search']), 'e2e_gcups', round(d['e2e']['value'],3), 'dp_only', round(d.get('dp_only_mode',{}).get('e2e_gcups',0),2))"
done
