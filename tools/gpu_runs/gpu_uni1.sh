#!/bin/bash
# warp-uniform search (default build) vs lane-0-only (-DRP_NO_UNIFORM): engine kernel time of 6000 config-2 families, oracle check of 2
PROBE_REPS=1 timeout 300 python tools/engine_probe.py 6000 2 > gpurun_out/uni1_default.log 2>&1; echo rc=$? >> gpurun_out/uni1_default.log
POASTA_B200_LIB=$PWD/variants/lib_nouni.so PROBE_REPS=1 timeout 300 python tools/engine_probe.py 6000 2 > gpurun_out/uni1_nouni.log 2>&1; echo rc=$? >> gpurun_out/uni1_nouni.log
( time timeout 600 python -m pytest tests/test_gpu_parity.py -x -q ) > gpurun_out/uni1_pytest.log 2>&1; echo rc=$? >> gpurun_out/uni1_pytest.log
