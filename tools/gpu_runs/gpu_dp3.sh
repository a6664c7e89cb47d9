#!/bin/bash
( time timeout 900 python -m pytest tests/test_gpu_dp_packed.py tests/test_gpu_parity.py -x -q --durations=5 ) > gpurun_out/dp3_pytest.log 2>&1; echo rc=$? >> gpurun_out/dp3_pytest.log
timeout 300 python tools/dp_kernel_probe.py 3552 10 > gpurun_out/dp3_probe_packed.log 2>&1; echo rc=$? >> gpurun_out/dp3_probe_packed.log
PROBE_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_dp_packed_launches.csv python tools/dp_kernel_probe.py 3552 10 > gpurun_out/dp3_launch.log 2>&1; echo rc=$? >> gpurun_out/dp3_launch.log
