#!/bin/bash
# After scripts/gpu_round1_final.sh has run under gpurun: turn the scratch reports in gpurun_out/ into the tracked
# summaries under profiles/ (runs here, no GPU needed).
set -e
for c in cfg2 cfg3 cfg4; do
  python scripts/ncu_to_profile.py gpurun_out/prof_final_$c.ncu-rep profiles/r01_k1_final_$c.md "round 1 final, K1 log-density kernel, $c" gpurun_out/plain_final_$c.log > /dev/null
done
python scripts/ncu_to_profile.py gpurun_out/prof_final_k3.ncu-rep profiles/r01_k3_final_cfg5.md "round 1 final, K3 score grid, configs[4] S=16384 F=10000 11x11" gpurun_out/plain_final_k3.log > /dev/null
cp gpurun_out/launches_final_cfg2.csv profiles/r01_launches_cfg2.csv
cp gpurun_out/bench_final.json profiles/r01_bench_final.json
python - <<'PY'
import subprocess, csv, io, json
out = {}
for name, rep in (("cfg2", "prof_final_cfg2"), ("cfg3", "prof_final_cfg3"), ("cfg4", "prof_final_cfg4"), ("k3_cfg5", "prof_final_k3")):
    txt = subprocess.run(["ncu", "-i", f"gpurun_out/{rep}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    d, u = dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))
    def get(k):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u[k], 1)
        return float(d[k].replace(",", "")) * mult
    r, w = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    out[name] = {"dram_bytes_read": r, "dram_bytes_write": w, "traffic": r + w, "ncu_duration_us": get("gpu__time_duration.sum")}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(out))
PY
