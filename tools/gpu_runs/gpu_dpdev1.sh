#!/bin/bash
( time POASTA_TIMING=1 timeout 900 python -m pytest tests/test_gpu_dp_device.py -x -q --durations=10 ) > gpurun_out/dpdev1_pytest.log 2>&1; echo rc=$? >> gpurun_out/dpdev1_pytest.log
( time timeout 300 python bench.py --config 2 --tiebreak dp --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/dpdev1_c2dp.json 2> gpurun_out/dpdev1_c2dp.err; echo rc=$? >> gpurun_out/dpdev1_c2dp.err
