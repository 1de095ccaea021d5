#!/bin/bash
# After scripts/gpu_round2_final.sh has run under gpurun: turn the scratch reports in gpurun_out/ into the tracked
# summaries under profiles/ (runs here, no GPU needed).
set -e
python scripts/ncu_to_profile.py gpurun_out/prof_r2_cfg3.ncu-rep profiles/r02_k1_cfg3.md "round 2, K1 log-density kernel, configs[2] (the bench headline)" gpurun_out/plain_r2_cfg3.log > /dev/null
python scripts/ncu_to_profile.py gpurun_out/prof_r2_cfg4.ncu-rep profiles/r02_k1d_cfg4.md "round 2, K1d (dynamic model, rewritten), configs[3]" gpurun_out/plain_r2_cfg4.log > /dev/null
python scripts/ncu_to_profile.py gpurun_out/prof_r2_k3.ncu-rep profiles/r02_k3_cfg5.md "round 2, K3 score grid (rewritten), configs[4] S=16384 F=10000 11x11" > /dev/null
python scripts/ncu_to_profile.py gpurun_out/prof_r2_nuts.ncu-rep profiles/r02_nuts_generic.md "round 2, NUTS step kernel with chain-minor state in steady state, configs[2] data, 32768 chains x 1339 parameters" > /dev/null
python scripts/ncu_to_profile.py gpurun_out/prof_r2_nuts_cm.ncu-rep profiles/r02_nuts_chain_major.md "round 2, NUTS step kernel with chain-major state (a warp per chain) in steady state, configs[2] data, 32768 chains x 1339 parameters" > /dev/null
python - <<'PY'
import json
for f, md in (("plain_r2_nuts.log", "profiles/r02_nuts_generic.md"), ("plain_r2_nuts_cm.log", "profiles/r02_nuts_chain_major.md")):
    d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
    s = open(md).read().split("\n")
    s.insert(3, "\nun-profiled run of the same command (CUDA events per block of 32 leapfrogs): leapfrog %.3f ms = log-density call %.3f ms + step %.3f ms (%s state)\n"
             % (d["ms_per_pair_steady"], d["k1_ms"], d["step_ms"], d["state_layout"]))
    open(md, "w").write("\n".join(s))
PY
python scripts/ncu_to_profile.py gpurun_out/prof_r2_few1.ncu-rep profiles/r02_k1_single_chain.md "round 2, K1 on ONE chain of configs[2] data (8-CTA cluster): what bounds the few-chain call" > /dev/null
cp gpurun_out/launches_r2_cfg3.csv profiles/r02_launches_cfg3.csv
cp gpurun_out/bench_r2_final.json profiles/r02_bench_final.json
cp gpurun_out/few_chain_r2.jsonl profiles/r02_few_chain.jsonl
for f in plain_r2_k3.log plain_r2_nuts.log plain_r2_nuts_cm.log; do cp gpurun_out/$f profiles/r02_${f#plain_r2_}; done
cp gpurun_out/layout_r2.jsonl profiles/r02_layouts.jsonl
cp gpurun_out/timeline_r2_cfg3.log profiles/r02_k1_timeline_cfg3.log
python - <<'PY'
import subprocess, csv, io, json
out = json.load(open("profiles/traffic.json"))
for name, rep in (("cfg3", "prof_r2_cfg3"), ("cfg4", "prof_r2_cfg4"), ("k3_cfg5", "prof_r2_k3"), ("nuts_generic_cfg3", "prof_r2_nuts"), ("nuts_chain_major_cfg3", "prof_r2_nuts_cm"),
                  ("cfg3_single_chain", "prof_r2_few1")):
    txt = subprocess.run(["ncu", "-i", f"gpurun_out/{rep}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    d, u = dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))
    def get(k):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u[k], 1)
        return float(d[k].replace(",", "")) * mult
    r, w = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    out[name] = {"dram_bytes_read": r, "dram_bytes_write": w, "traffic": r + w, "ncu_duration_us": get("gpu__time_duration.sum"),
                 "round": 2}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if v.get("round") == 2}))
PY
