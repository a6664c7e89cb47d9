#!/bin/bash
( time timeout 600 python -m pytest tests/test_gpu_dp_packed.py -x -q --durations=6 ) > gpurun_out/dp5_pytest.log 2>&1; echo rc=$? >> gpurun_out/dp5_pytest.log
( time timeout 300 python bench.py --config 4 --tiebreak dp --steps 1 --warmup 1 --no-cpu-baseline ) > gpurun_out/dp5_c4dp.json 2> gpurun_out/dp5_c4dp.err; echo rc=$? >> gpurun_out/dp5_c4dp.err
( time timeout 300 python bench.py --config 2 --tiebreak dp --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/dp5_c2dp.json 2> gpurun_out/dp5_c2dp.err; echo rc=$? >> gpurun_out/dp5_c2dp.err
