#!/bin/bash
( time timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py tests/test_cli.py "tests/test_gpu_engine.py::test_config5_full_size_graph_reads_sample" -x -q -m gpu ) > gpurun_out/blk3_pytest.log 2>&1; echo rc=$? >> gpurun_out/blk3_pytest.log
( POASTA_TIMING=1 timeout 600 python bench.py --config 5 --steps 3 --warmup 3 ) > gpurun_out/blk3_c5.json 2> gpurun_out/blk3_c5.err; echo rc=$? >> gpurun_out/blk3_c5.err
