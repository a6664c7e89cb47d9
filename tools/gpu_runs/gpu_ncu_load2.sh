#!/bin/bash
# stall picture of the engine kernel under full load (32 workers per SM): sampling sections only, no SASS patching
timeout 600 ncu --section WarpStateStats --section SchedulerStats --section SpeedOfLight --section MemoryWorkloadAnalysis --clock-control none -k regex:msa_engine -c 1 -o gpurun_out/r02_engine_load -f python tools/ncu_engine_probe.py 4736 > gpurun_out/ncu_load2.log 2>&1; echo rc=$? >> gpurun_out/ncu_load2.log
