#!/bin/bash
# the driver's launch of the scaling bench at N = 8 (shorter: 2 timed steps), then the reference arm
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 2 --warmup 2 ) > gpurun_out/n8_bench.json 2> gpurun_out/n8_bench.err; echo rc=$? >> gpurun_out/n8_bench.err
nproc >> gpurun_out/n8_bench.err; free -g | head -2 >> gpurun_out/n8_bench.err
