#!/bin/bash
# one gpurun call: DP chunk test, bench lines of configs 3 / 4 / 5 (one step each), ncu --set full capture of the engine kernel
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunked or dp_mode" > gpurun_out/r02_pytest3.log 2>&1; echo rc=$? >> gpurun_out/r02_pytest3.log
for c in 3 4 5; do
  timeout 500 python bench.py --config $c --steps 1 --warmup 1 > gpurun_out/r02_bench_c$c.json 2> gpurun_out/r02_bench_c$c.err; echo rc=$? >> gpurun_out/r02_bench_c$c.err
done
python tools/ncu_engine_probe.py > gpurun_out/ncu_probe_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:msa_engine_kernel -c 1 -o gpurun_out/r02_engine --force-overwrite python tools/ncu_engine_probe.py > gpurun_out/ncu_engine.log 2>&1; echo rc=$? >> gpurun_out/ncu_engine.log
