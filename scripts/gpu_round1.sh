set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 600 gpurun_out/bench_r1.err
A="--steps 5 --warmup 3 --no-extras --no-graph --cpu-seconds 0.5"
python bench.py $A > gpurun_out/plain_cfg2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py $A > gpurun_out/ncu_cfg2.log 2>&1
python bench.py $A > gpurun_out/plain_cfg2b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:logdensity -s 4 -c 2 -o gpurun_out/prof_cfg2 python bench.py $A > gpurun_out/ncu_cfg2_full.log 2>&1
B="--workload cfg3 --steps 2 --warmup 3 --no-extras --no-graph --cpu-seconds 0.5 --cpu-chains 8"
python bench.py $B > gpurun_out/plain_cfg3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:logdensity -s 4 -c 1 -o gpurun_out/prof_cfg3 python bench.py $B > gpurun_out/ncu_cfg3_full.log 2>&1
tail -2 gpurun_out/plain_cfg3.log | cut -c1-1500
ls -la gpurun_out
