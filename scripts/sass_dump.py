"""Dump SASS [lo,hi) of the first kernel in an ncu report with executed counts and samples."""
import csv, io, subprocess, sys
rep, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr=None; data=[]; nk=0
for r in rows:
    if r and r[0]=="Kernel Name":
        nk+=1
        if nk>1: break
        continue
    if r and r[0]=="Address": hdr=r; continue
    if hdr and len(r)==len(hdr): data.append(r)
ia, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
mx = max(int(r[ia]) for r in data[lo:hi])
for i in range(lo, min(hi, len(data))):
    r = data[i]
    if int(r[ia]) >= thr * mx: print(i, r[ia].rjust(10), r[ismp].rjust(6), r[isrc].strip()[:100])
