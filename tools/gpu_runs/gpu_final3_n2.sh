#!/bin/bash
# the driver's launch line at N = 2 (short: 2 timed steps) with the final build
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 2 ) > gpurun_out/final3_n2.json 2> gpurun_out/final3_n2.err; echo rc=$? >> gpurun_out/final3_n2.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 3 --tiebreak dp ) > gpurun_out/final3_n2_dp.json 2> gpurun_out/final3_n2_dp.err; echo rc=$? >> gpurun_out/final3_n2_dp.err
