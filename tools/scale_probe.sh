#!/bin/bash
# Search-kernel latency vs alignments in flight (one pipeline lane, rounds 1..7 of config-2 families).
for n in 148 592 1184 2368 4736 9472; do
  POASTA_LANES=1 POASTA_TIMING=1 python bench.py --families $n --steps 1 --warmup 1 --seqs 8 --no-cpu-baseline 2> gpurun_out/probe_$n.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print($n, d['ms_per_step'], d['device_ms_per_step'], d['value'])"
  grep lane gpurun_out/probe_$n.err | tail -2
done
