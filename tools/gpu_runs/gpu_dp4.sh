#!/bin/bash
( time timeout 600 python -m pytest tests/test_gpu_dp_packed.py tests/test_gpu_parity.py -x -q --durations=6 ) > gpurun_out/dp4_pytest.log 2>&1; echo rc=$? >> gpurun_out/dp4_pytest.log
