"""Per-CUDA-source-line executed-instruction and stall-sample shares from an ncu report (cuda,sass view)."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; fname = ""; per = []; nfn = 0; fn = None
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Function Name":
        if fn is None: fn = r[1]
        if r[1] != fn: break
        continue
    if r and r[0] == "Line No": hdr = r; continue
    if not hdr or len(r) != len(hdr) or not r[0].isdigit(): continue
    ie = hdr.index("Instructions Executed"); ism = hdr.index("# Samples")
    per.append((fname, int(r[0]), r[1].strip(), int(r[ie] or 0), int(r[ism] or 0)))
tot = sum(p[3] for p in per) or 1; ts = sum(p[4] for p in per) or 1
print(fn); print("total executed", tot, "samples", ts)
top = sorted(per, key=lambda p: -(p[3] / tot + p[4] / ts))[:N]
for f, ln, src, e, s in sorted(top, key=lambda p: (p[0], p[1])):
    print(f"{f[:14]:14s}{ln:>5} exec {100*e/tot:5.1f}%  smp {100*s/ts:5.1f}%  {src[:95]}")
