#!/bin/bash
( time timeout 460 python -m pytest tests/test_gpu_engine.py tests/test_gpu_dp_device.py -x -q -m gpu ) > gpurun_out/final8_pytest.log 2>&1; echo rc=$? >> gpurun_out/final8_pytest.log
