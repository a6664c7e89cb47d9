#!/bin/bash
( time timeout 120 python bench.py --tiebreak dp --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/final10_bench_dp.json 2> gpurun_out/final10_bench_dp.err; echo rc=$? >> gpurun_out/final10_bench_dp.err
