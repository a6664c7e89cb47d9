#!/bin/bash
( time timeout 900 python -m pytest tests/test_gpu_dp_device.py tests/test_gpu_parity.py -x -q ) > gpurun_out/ovl1_pytest.log 2>&1; echo rc=$? >> gpurun_out/ovl1_pytest.log
( POASTA_TIMING=1 timeout 300 python tools/msa_phase_probe.py 6000 dp ) 2>&1 | grep -v "dp round" > gpurun_out/ovl1_probe.log
( POASTA_NO_OVERLAP=1 POASTA_TIMING=1 timeout 300 python tools/msa_phase_probe.py 6000 dp ) 2>&1 | grep -v "dp round" > gpurun_out/ovl1_probe_serial.log
