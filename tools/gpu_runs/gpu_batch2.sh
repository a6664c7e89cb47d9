#!/bin/bash
# full GPU test suite + default bench (config 2) with both arms
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r02_pytest_full.log 2>&1; echo rc=$? >> gpurun_out/r02_pytest_full.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo rc=$? >> gpurun_out/r02_bench_ref.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err; echo rc=$? >> gpurun_out/r02_bench_c2.err
