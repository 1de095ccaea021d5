"""Warp-stall samples and executed instructions between consecutive barriers (BAR.SYNC) of the first kernel."""
import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; data = []; nk = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        nk += 1
        if nk > 1: break
        continue
    if r and r[0] == "Address": hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
ia, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[ia]) for r in data); ts = sum(int(r[ismp]) for r in data)
print("sass instrs", len(data), "executed", tot, "samples", ts)
start = 0; e = s = 0; k = 0
for i, r in enumerate(data):
    e += int(r[ia]); s += int(r[ismp])
    if "BAR.SYNC" in r[isrc] or "EXIT" in r[isrc] and int(r[ia]) > 0 or i == len(data) - 1:
        if e * 500 > tot or s * 200 > ts:
            print(f"region {k}: sass [{start},{i}] exec {100*e/tot:5.1f}%  samples {100*s/ts:5.1f}%  ends with {r[isrc].strip()[:40]}")
        k += 1; start = i + 1; e = s = 0
