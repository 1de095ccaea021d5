#!/bin/bash
# Round-2 evidence run A (under gpurun, one GPU): K1d capture, fit-step timing + one capture of the generic NUTS step kernel.
mkdir -p gpurun_out
Cc="--workload cfg4 --radius 0.5 --steps 2 --warmup 3 --no-extras --no-subrecords --no-graph --cpu-seconds 0.2 --cpu-chains 8"
timeout 300 python bench.py $Cc > gpurun_out/plain_r2_k1d.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:logdensity -s 4 -c 1 -f -o gpurun_out/prof_r2_k1d python bench.py $Cc > gpurun_out/ncu_r2_k1d.log 2>&1
tail -c 300 gpurun_out/plain_r2_k1d.log
timeout 300 python scripts/fit_step_time.py 32768 96 > gpurun_out/fit_step_32768.json 2> gpurun_out/fit_step_32768.err; cat gpurun_out/fit_step_32768.json; tail -3 gpurun_out/fit_step_32768.err
timeout 300 python scripts/fit_step_time.py 4096 192 > gpurun_out/fit_step_4096.json 2> gpurun_out/fit_step_4096.err; cat gpurun_out/fit_step_4096.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nuts_step_kernel -s 40 -c 1 -f -o gpurun_out/prof_r2_nuts_generic python scripts/fit_step_time.py 32768 64 > gpurun_out/ncu_r2_nuts.log 2>&1
ls -la gpurun_out | grep r2_
