#!/bin/bash
# pool-backed graph lists (host side only): smoke, the graph-heavy GPU tests, host phases of a DP build
( time timeout 60 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/final9_smoke.log 2>&1; echo rc=$? >> gpurun_out/final9_smoke.log
( time timeout 100 python -m pytest tests/test_gpu_parity.py "tests/test_gpu_engine.py::test_ragged_families_and_counters" "tests/test_gpu_engine.py::test_existing_graphs_continue_on_the_device" "tests/test_gpu_dp_device.py::test_existing_graphs_continue_on_the_device" "tests/test_gpu_dp_device.py::test_ragged_families_match_the_dense_model_build" -x -q ) > gpurun_out/final9_pytest.log 2>&1; echo rc=$? >> gpurun_out/final9_pytest.log
( POASTA_TIMING=1 timeout 60 python tools/msa_phase_probe.py 6000 dp ) 2>&1 | grep -v "dp round" > gpurun_out/final9_probe.log
