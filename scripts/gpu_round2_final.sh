#!/bin/bash
# Round-2 evidence run (under gpurun, one GPU): full bench line, launch list, ncu --set full captures of K1 (configs[2]), K3,
# K1d (configs[3]), the NUTS step in steady state with chain-minor and chain-major state, and the single-chain K1 call.  Every ncu run comes directly after
# the same command exited 0 without ncu.  scripts/refresh_profiles_r2.sh turns gpurun_out/ into profiles/.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; echo "bench rc=$?"
A="--steps 3 --warmup 3 --no-extras --no-graph --cpu-seconds 0.2 --cpu-chains 8 --fit-warmup 2 --fit-samples 2 --fit-chains 4096"
python bench.py $A > gpurun_out/plain_r2_launches.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2_cfg3.csv python bench.py $A > gpurun_out/ncu_r2_launches.log 2>&1
B="--steps 2 --warmup 3 --no-extras --no-subrecords --no-graph --cpu-seconds 0.2 --cpu-chains 8"
python bench.py $B > gpurun_out/plain_r2_cfg3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:logdensity_kernel -s 4 -c 1 -f -o gpurun_out/prof_r2_cfg3 python bench.py $B > gpurun_out/ncu_r2_cfg3.log 2>&1
Cc="--workload cfg4 --radius 0.5 --steps 2 --warmup 3 --no-extras --no-subrecords --no-graph --cpu-seconds 0.2 --cpu-chains 8"
python bench.py $Cc > gpurun_out/plain_r2_cfg4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:logdensity_dynamic -s 4 -c 1 -f -o gpurun_out/prof_r2_cfg4 python bench.py $Cc > gpurun_out/ncu_r2_cfg4.log 2>&1
python scripts/grid_bench.py > gpurun_out/plain_r2_k3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:score_grid_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_k3 python scripts/grid_bench.py > gpurun_out/ncu_r2_k3.log 2>&1
export BPLX_NO_GRAPH=1  # (plain launches, so that ncu's launch skip count addresses a steady-state step)
python scripts/fit_step_time.py 32768 384 0 chain_minor > gpurun_out/plain_r2_nuts.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nuts_step_kernel -s 300 -c 1 -f -o gpurun_out/prof_r2_nuts python scripts/fit_step_time.py 32768 384 0 chain_minor > gpurun_out/ncu_r2_nuts.log 2>&1
python scripts/fit_step_time.py 32768 384 0 chain_major > gpurun_out/plain_r2_nuts_cm.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nuts_step_kernel -s 300 -c 1 -f -o gpurun_out/prof_r2_nuts_cm python scripts/fit_step_time.py 32768 384 0 chain_major > gpurun_out/ncu_r2_nuts_cm.log 2>&1
unset BPLX_NO_GRAPH
python scripts/layout_time.py cfg3 > gpurun_out/layout_r2.jsonl 2>&1
python scripts/timeline.py cfg3 2.0 > gpurun_out/timeline_r2_cfg3.log 2>&1  # (needs scripts/build_timeline.py run beforehand)
python scripts/few_chain_time.py 1 > gpurun_out/plain_r2_few1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:logdensity_kernel -s 10 -c 1 -f -o gpurun_out/prof_r2_few1 python scripts/few_chain_time.py 1 > gpurun_out/ncu_r2_few1.log 2>&1
python scripts/few_chain_time.py > gpurun_out/few_chain_r2.jsonl 2>&1
ls -la gpurun_out | grep r2_
