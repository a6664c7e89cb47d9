#!/bin/bash
# the driver's own N = 1 commands: reference arm, then the repo arm, 20 timed steps after 5 warm-up steps
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/final4_ref.json 2> gpurun_out/final4_ref.err; echo rc=$? >> gpurun_out/final4_ref.err
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/final4_bench.json 2> gpurun_out/final4_bench.err; echo rc=$? >> gpurun_out/final4_bench.err
