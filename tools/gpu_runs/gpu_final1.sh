#!/bin/bash
# full GPU suite, smoke, and the default bench line (exact mode) with the current build
( time timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 ) > gpurun_out/final1_pytest.log 2>&1; echo rc=$? >> gpurun_out/final1_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/final1_smoke.log 2>&1; echo rc=$? >> gpurun_out/final1_smoke.log
( time timeout 600 python bench.py --steps 3 --warmup 3 ) > gpurun_out/final1_bench.json 2> gpurun_out/final1_bench.err; echo rc=$? >> gpurun_out/final1_bench.err
