#!/bin/bash
( POASTA_TIMING=1 timeout 600 python bench.py --config 5 --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/blk4_c5.json 2> gpurun_out/blk4_c5.err; echo rc=$? >> gpurun_out/blk4_c5.err
