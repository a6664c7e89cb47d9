#!/bin/bash
( time timeout 60 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/final11_smoke.log 2>&1; echo rc=$? >> gpurun_out/final11_smoke.log
( time timeout 40 python -m pytest tests/test_gpu_parity.py -x -q ) > gpurun_out/final11_pytest.log 2>&1; echo rc=$? >> gpurun_out/final11_pytest.log
