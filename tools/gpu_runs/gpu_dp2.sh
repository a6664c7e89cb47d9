#!/bin/bash
# launch list + full ncu capture of the packed DP fill kernel (after the same command ran clean without ncu in gpu_dp1.sh)
PROBE_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_dp_packed_launches.csv python tools/dp_kernel_probe.py 3552 10 > gpurun_out/dp2_launch.log 2>&1; echo rc=$? >> gpurun_out/dp2_launch.log
PROBE_REPS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:dp_fill_packed -c 1 -o gpurun_out/r02_dp_packed -f python tools/dp_kernel_probe.py 3552 10 > gpurun_out/dp2_ncu.log 2>&1; echo rc=$? >> gpurun_out/dp2_ncu.log
PROBE_REPS=1 timeout 900 ncu --set full --clock-control none -k regex:traceback_rows -c 1 -o gpurun_out/r02_tb_rows -f python tools/dp_kernel_probe.py 3552 10 > gpurun_out/dp2_ncu_tb.log 2>&1; echo rc=$? >> gpurun_out/dp2_ncu_tb.log
