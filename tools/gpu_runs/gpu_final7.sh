#!/bin/bash
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dp_packed.py tests/test_gpu_configs.py tests/test_cli.py -x -q -m gpu ) > gpurun_out/final7_pytest.log 2>&1; echo rc=$? >> gpurun_out/final7_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/final7_smoke.log 2>&1; echo rc=$? >> gpurun_out/final7_smoke.log
( POASTA_TIMING=1 timeout 600 python bench.py --config 5 --steps 5 --warmup 3 ) > gpurun_out/final7_c5.json 2> gpurun_out/final7_c5.err; echo rc=$? >> gpurun_out/final7_c5.err
