#!/bin/bash
# A/B of the search-core change: engine + parity GPU tests, then 6000 config-2 families through the engine (2 reps), oracle check of 8
( time timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_parity.py -x -q --durations=5 ) > gpurun_out/ab1_pytest.log 2>&1; echo rc=$? >> gpurun_out/ab1_pytest.log
timeout 600 python tools/engine_probe.py 6000 8 > gpurun_out/ab1_probe6000.log 2>&1; echo rc=$? >> gpurun_out/ab1_probe6000.log
