#!/bin/bash
# launch list of the default bench command (reference-exact mode), after the same command exited 0 without ncu
timeout 600 python bench.py --steps 2 --warmup 3 --families 1184 --no-cpu-baseline > gpurun_out/final5_plain.json 2> gpurun_out/final5_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench_exact_1184families.csv python bench.py --steps 2 --warmup 3 --families 1184 --no-cpu-baseline > gpurun_out/final5_ncu.json 2> gpurun_out/final5_ncu.err; echo rc=$? >> gpurun_out/final5_ncu.err
