#!/bin/bash
# After scripts/gpu_round2_final.sh has run under gpurun: turn the scratch reports in gpurun_out/ into the tracked
# summaries under profiles/ (runs here, no GPU needed).
set -e
python scripts/ncu_to_profile.py gpurun_out/prof_r2_cfg3.ncu-rep profiles/r02_k1_cfg3.md "round 2, K1 log-density kernel, configs[2] (the bench headline)" gpurun_out/plain_r2_cfg3.log > /dev/null
python scripts/ncu_to_profile.py gpurun_out/prof_r2_cfg4.ncu-rep profiles/r02_k1d_cfg4.md "round 2, K1d (dynamic model, rewritten), configs[3]" gpurun_out/plain_r2_cfg4.log > /dev/null
python scripts/ncu_to_profile.py gpurun_out/prof_r2_k3.ncu-rep profiles/r02_k3_cfg5.md "round 2, K3 score grid (rewritten), configs[4] S=16384 F=10000 11x11" > /dev/null
python scripts/ncu_to_profile.py gpurun_out/prof_r2_nuts.ncu-rep profiles/r02_nuts_generic.md "round 2, generic NUTS step kernel in steady state, configs[2] data, 32768 chains x 1339 parameters" > /dev/null
python scripts/ncu_to_profile.py gpurun_out/prof_r2_few1.ncu-rep profiles/r02_k1_single_chain.md "round 2, K1 on ONE chain of configs[2] data (8-CTA cluster): what bounds the few-chain call" > /dev/null
cp gpurun_out/launches_r2_cfg3.csv profiles/r02_launches_cfg3.csv
cp gpurun_out/bench_r2_final.json profiles/r02_bench_final.json
cp gpurun_out/few_chain_r2.jsonl profiles/r02_few_chain.jsonl
for f in plain_r2_k3.log plain_r2_nuts.log; do cp gpurun_out/$f profiles/r02_${f#plain_r2_}; done
python - <<'PY'
import subprocess, csv, io, json
out = json.load(open("profiles/traffic.json"))
for name, rep in (("cfg3", "prof_r2_cfg3"), ("cfg4", "prof_r2_cfg4"), ("k3_cfg5", "prof_r2_k3"), ("nuts_generic_cfg3", "prof_r2_nuts"),
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
