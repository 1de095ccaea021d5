#!/bin/bash
# usage: bash scripts/gpu_check.sh <tag>   (run under gpurun): GPU tests + a full bench line
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_${TAG}.log
python bench.py --cpu-seconds 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_${TAG}.err
python - <<'PY' gpurun_out/bench_${TAG}.json
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4g ms/step %.5f e2e %.4g frac %.3f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["frac"]))
for e in j.get("extra_workloads", []): print(e["workload"][:40], e["value"], e.get("fp32_frac"))
PY
