#!/bin/bash
# usage: bash scripts/gpu_quick.sh <tag>  -- timing only (cfg2 + extras), no tests, tiny CPU leg
TAG=${1:-x}
mkdir -p gpurun_out
python bench.py --cpu-seconds 0.5 --cpu-chains 32 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_${TAG}.err
python - <<'PY' gpurun_out/bench_${TAG}.json
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("cfg2 value %.4g us/step %.2f e2e %.4g frac %.3f" % (j["value"], 1e3*j["ms_per_step"], j["e2e"]["value"], j["roofline"]["frac"]))
for e in j.get("extra_workloads", []): print(e["workload"][:40], "%.4g" % e["value"], e.get("ms_per_call"), e.get("fp32_frac"))
PY
