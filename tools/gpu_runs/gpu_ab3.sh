#!/bin/bash
# L1 / shared-memory carve-out and queue-window size of the engine kernel: kernel time of 6000 config-2 families
run() { # name lib carveout
  POASTA_B200_LIB=$PWD/variants/lib_$2.so PROBE_REPS=1 ${3:+POASTA_ENGINE_CARVEOUT=$3} timeout 300 python tools/engine_probe.py 6000 2 > gpurun_out/ab3_$1.log 2>&1; echo rc=$? >> gpurun_out/ab3_$1.log
}
run h0_default h0
POASTA_ENGINE_CARVEOUT=29 run h0_c29 h0
POASTA_ENGINE_CARVEOUT=45 run h0_c45 h0
POASTA_ENGINE_CARVEOUT=100 run h0_c100 h0
POASTA_ENGINE_CARVEOUT=29 run q32_c29 q32
POASTA_ENGINE_CARVEOUT=15 run q16_c15 q16
