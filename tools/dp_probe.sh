#!/bin/bash
# DP-only mode, one pipeline lane (clean per-kernel times), under library variants: tools/dp_probe.sh v5 v6 ...
for v in "$@"; do
  POASTA_B200_LIB=$PWD/variants/lib_$v.so POASTA_LANES=1 python bench.py --families 1776 --seqs 12 --steps 1 --warmup 1 --tiebreak dp --no-cpu-baseline 2>gpurun_out/dp_probe_$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$v', 'step_ms', round(d['ms_per_step']), 'dp_ms', round(d['device_ms_per_step']['dp_fill+traceback'],1), 'kernel_gcups', round(d['value'],2))"
done
