#!/bin/bash
for c in 2 3 4; do
( time timeout 400 python bench.py --config $c --tiebreak dp --steps 3 --warmup 3 ) > gpurun_out/dpdev3_c${c}dp.json 2> gpurun_out/dpdev3_c${c}dp.err; echo rc=$? >> gpurun_out/dpdev3_c${c}dp.err
done
( POASTA_TIMING=1 timeout 300 python tools/msa_phase_probe.py 6000 dp ) 2>&1 | grep -v "dp round" > gpurun_out/dpdev3_probe.log
