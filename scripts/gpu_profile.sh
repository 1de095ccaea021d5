#!/bin/bash
# usage: bash scripts/gpu_profile.sh <tag>   (run under gpurun; writes gpurun_out/prof_<tag>_cfg{2,3}.ncu-rep)
TAG=${1:-x}
mkdir -p gpurun_out
A="--steps 5 --warmup 3 --no-extras --no-graph --cpu-seconds 0.2 --cpu-chains 8"
python bench.py $A > gpurun_out/plain_${TAG}_cfg2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:logdensity -s 4 -c 1 -o gpurun_out/prof_${TAG}_cfg2 python bench.py $A > gpurun_out/ncu_${TAG}_cfg2.log 2>&1
B="--workload cfg3 --steps 2 --warmup 3 --no-extras --no-graph --cpu-seconds 0.2 --cpu-chains 8"
python bench.py $B > gpurun_out/plain_${TAG}_cfg3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:logdensity -s 4 -c 1 -o gpurun_out/prof_${TAG}_cfg3 python bench.py $B > gpurun_out/ncu_${TAG}_cfg3.log 2>&1
ls -la gpurun_out | tail -8
