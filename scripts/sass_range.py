"""Print the SASS of the first kernel in an ncu report between two instruction indices, with executed counts and samples.
usage: python scripts/sass_range.py rep.ncu-rep START END [min_exec]"""
import csv, io, subprocess, sys
rep, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
mn = int(sys.argv[4]) if len(sys.argv) > 4 else 0
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
ia, isrc, ismp, iad = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Address")
for i in range(a, min(b + 1, len(data))):
    r = data[i]
    if int(r[ia]) >= mn:
        print(f"{i:5d} {r[iad][-5:]} {int(r[ia]):8d} {int(r[ismp]):4d}  {r[isrc].strip()[:100]}")
