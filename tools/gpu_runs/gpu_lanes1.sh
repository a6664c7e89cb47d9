#!/bin/bash
( time timeout 900 python -m pytest tests/test_gpu_dp_device.py -x -q ) > gpurun_out/lanes1_pytest.log 2>&1; echo rc=$? >> gpurun_out/lanes1_pytest.log
for l in 3 2 1 4; do
( POASTA_DP_LANES=$l POASTA_TIMING=1 timeout 300 python tools/msa_phase_probe.py 6000 dp ) 2>&1 | grep -v "dp round" > gpurun_out/lanes1_probe_$l.log
done
