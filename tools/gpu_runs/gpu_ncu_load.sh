#!/bin/bash
# ncu of the engine kernel under full load: 4736 config-2 families cut to 4 records (32 workers per SM), source-level sampling
timeout 300 python tools/ncu_engine_probe.py 4736 > gpurun_out/ncu_load_plain.log 2>&1; echo rc=$? >> gpurun_out/ncu_load_plain.log
timeout 900 ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section SourceCounters --section MemoryWorkloadAnalysis --section InstructionStats --section LaunchStats --section Occupancy \
  --import-source on --clock-control none -k regex:msa_engine -c 1 -o gpurun_out/r02_engine_load -f python tools/ncu_engine_probe.py 4736 > gpurun_out/ncu_load.log 2>&1; echo rc=$? >> gpurun_out/ncu_load.log
