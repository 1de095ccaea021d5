#!/bin/bash
# Round-1 evidence run (under gpurun, one GPU): full bench line, launch list, ncu --set full captures of K1 (cfg2, cfg3),
# K1d (cfg4) and K3.  Every ncu run comes directly after the same command exited 0 without ncu.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
A="--steps 5 --warmup 3 --no-extras --no-graph --cpu-seconds 0.2 --cpu-chains 8"
python bench.py $A > gpurun_out/plain_final_cfg2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_final_cfg2.csv python bench.py $A > gpurun_out/ncu_launches.log 2>&1
python bench.py $A > gpurun_out/plain_final_cfg2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:logdensity -s 4 -c 1 -o gpurun_out/prof_final_cfg2 python bench.py $A > gpurun_out/ncu_final_cfg2.log 2>&1
B="--workload cfg3 --steps 2 --warmup 3 --no-extras --no-graph --cpu-seconds 0.2 --cpu-chains 8"
python bench.py $B > gpurun_out/plain_final_cfg3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:logdensity -s 4 -c 1 -o gpurun_out/prof_final_cfg3 python bench.py $B > gpurun_out/ncu_final_cfg3.log 2>&1
Cc="--workload cfg4 --radius 0.5 --steps 2 --warmup 3 --no-extras --no-graph --cpu-seconds 0.2 --cpu-chains 8"
python bench.py $Cc > gpurun_out/plain_final_cfg4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:logdensity -s 4 -c 1 -o gpurun_out/prof_final_cfg4 python bench.py $Cc > gpurun_out/ncu_final_cfg4.log 2>&1
python scripts/grid_bench.py > gpurun_out/plain_final_k3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:score_grid_kernel -s 3 -c 1 -o gpurun_out/prof_final_k3 python scripts/grid_bench.py > gpurun_out/ncu_final_k3.log 2>&1
ls -la gpurun_out | grep final
