#!/bin/bash
( POASTA_TIMING=1 timeout 300 python tools/msa_phase_probe.py 6000 dp ) > gpurun_out/dpdev2_probe.log 2>&1; echo rc=$? >> gpurun_out/dpdev2_probe.log
nproc >> gpurun_out/dpdev2_probe.log
