#!/bin/bash
# launch list of the device-resident DP rounds (one 6,000-family config-2 batch), after the same command ran clean without ncu
PROBE_REPS=1 timeout 300 python tools/msa_phase_probe.py 6000 dp > gpurun_out/dpdev4_plain.log 2>&1 || exit 1
PROBE_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_dp_device_rounds.csv python tools/msa_phase_probe.py 6000 dp > gpurun_out/dpdev4_ncu.log 2>&1; echo rc=$? >> gpurun_out/dpdev4_ncu.log
