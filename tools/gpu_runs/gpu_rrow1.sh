#!/bin/bash
# reached-score rows (default build) vs reached-offset bitmap (variants/lib_base.so): engine kernel time of 6000 config-2 families
PROBE_REPS=1 timeout 300 python tools/engine_probe.py 6000 2 > gpurun_out/rrow1_new.log 2>&1; echo rc=$? >> gpurun_out/rrow1_new.log
POASTA_B200_LIB=$PWD/variants/lib_base.so PROBE_REPS=1 timeout 300 python tools/engine_probe.py 6000 2 > gpurun_out/rrow1_base.log 2>&1; echo rc=$? >> gpurun_out/rrow1_base.log
( time timeout 600 python -m pytest tests/test_gpu_parity.py "tests/test_gpu_engine.py::test_ragged_families_and_counters" -x -q ) > gpurun_out/rrow1_pytest.log 2>&1; echo rc=$? >> gpurun_out/rrow1_pytest.log
