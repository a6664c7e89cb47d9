#!/bin/bash
( time timeout 900 python -m pytest tests/test_gpu_dp_device.py tests/test_gpu_dp_packed.py tests/test_cli.py -x -q -m gpu ) > gpurun_out/final2_pytest.log 2>&1; echo rc=$? >> gpurun_out/final2_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/final2_smoke.log 2>&1; echo rc=$? >> gpurun_out/final2_smoke.log
( time timeout 400 python bench.py --tiebreak dp --steps 5 --warmup 3 ) > gpurun_out/final2_bench_dp.json 2> gpurun_out/final2_bench_dp.err; echo rc=$? >> gpurun_out/final2_bench_dp.err
( time timeout 400 python bench.py --config 3 --tiebreak dp --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/final2_bench_dp_c3.json 2> gpurun_out/final2_bench_dp_c3.err; echo rc=$? >> gpurun_out/final2_bench_dp_c3.err
