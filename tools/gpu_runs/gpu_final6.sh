#!/bin/bash
# final state of the round: full GPU suite, smoke, config-5 and default bench lines
( time timeout 1500 python -m pytest tests -x -q -m gpu --durations=6 ) > gpurun_out/final6_pytest.log 2>&1; echo rc=$? >> gpurun_out/final6_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/final6_smoke.log 2>&1; echo rc=$? >> gpurun_out/final6_smoke.log
( time timeout 600 python bench.py --config 5 --steps 5 --warmup 3 ) > gpurun_out/final6_c5.json 2> gpurun_out/final6_c5.err; echo rc=$? >> gpurun_out/final6_c5.err
( time timeout 600 python bench.py --steps 2 --warmup 3 ) > gpurun_out/final6_bench.json 2> gpurun_out/final6_bench.err; echo rc=$? >> gpurun_out/final6_bench.err
