#!/bin/bash
for v in uni_only uni_gather uni_prune; do
POASTA_B200_LIB=$PWD/variants/lib_$v.so PROBE_REPS=1 timeout 300 python tools/engine_probe.py 6000 2 > gpurun_out/uni2_$v.log 2>&1; echo rc=$? >> gpurun_out/uni2_$v.log
done
