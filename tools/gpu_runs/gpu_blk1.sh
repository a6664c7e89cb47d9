#!/bin/bash
( time timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py "tests/test_gpu_engine.py::test_config5_full_size_graph_reads_sample" -x -q ) > gpurun_out/blk1_pytest.log 2>&1; echo rc=$? >> gpurun_out/blk1_pytest.log
( POASTA_TIMING=1 timeout 600 python bench.py --config 5 --steps 2 --warmup 3 ) > gpurun_out/blk1_c5.json 2> gpurun_out/blk1_c5.err; echo rc=$? >> gpurun_out/blk1_c5.err
( POASTA_BLOCKED_TABLE=0 POASTA_TIMING=1 timeout 600 python bench.py --config 5 --steps 1 --warmup 1 --no-cpu-baseline ) > gpurun_out/blk1_c5_dense.json 2> gpurun_out/blk1_c5_dense.err; echo rc=$? >> gpurun_out/blk1_c5_dense.err
PROBE_REPS=1 timeout 300 python tools/engine_probe.py 6000 2 > gpurun_out/blk1_engine.log 2>&1; echo rc=$? >> gpurun_out/blk1_engine.log
