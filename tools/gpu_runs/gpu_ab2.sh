#!/bin/bash
# A/B of search-core variants (tools/build_variant.sh): kernel time of 6000 config-2 families through the engine, oracle check of 3
for v in "$@"; do
  POASTA_B200_LIB=$PWD/variants/lib_$v.so PROBE_REPS=1 timeout 300 python tools/engine_probe.py 6000 3 > gpurun_out/ab2_$v.log 2>&1; echo rc=$? >> gpurun_out/ab2_$v.log
done
