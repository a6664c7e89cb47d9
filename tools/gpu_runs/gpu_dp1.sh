#!/bin/bash
( time timeout 900 python -m pytest tests/test_gpu_dp_packed.py -x -q --durations=5 ) > gpurun_out/dp1_pytest.log 2>&1; echo rc=$? >> gpurun_out/dp1_pytest.log
timeout 300 python tools/dp_kernel_probe.py 3552 10 > gpurun_out/dp1_probe_packed.log 2>&1; echo rc=$? >> gpurun_out/dp1_probe_packed.log
POASTA_DP_KERNEL=warp32 timeout 300 python tools/dp_kernel_probe.py 3552 10 > gpurun_out/dp1_probe_warp32.log 2>&1; echo rc=$? >> gpurun_out/dp1_probe_warp32.log
