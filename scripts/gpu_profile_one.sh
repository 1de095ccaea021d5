#!/bin/bash
# usage: bash scripts/gpu_profile_one.sh <tag> <bench args...>   (run under gpurun; one ncu --set full capture of K1)
TAG=$1; shift
mkdir -p gpurun_out
A="$@ --steps 3 --warmup 3 --no-extras --no-graph --cpu-seconds 0.2 --cpu-chains 8"
python bench.py $A > gpurun_out/plain_${TAG}.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:logdensity -s 4 -c 1 -o gpurun_out/prof_${TAG} python bench.py $A > gpurun_out/ncu_${TAG}.log 2>&1
tail -c 300 gpurun_out/plain_${TAG}.log; ls -la gpurun_out | grep ${TAG}
