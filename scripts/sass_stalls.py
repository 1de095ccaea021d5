"""Top stalled SASS instructions + stall-reason totals from an ncu sass source-page CSV (first kernel)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
data = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name':
        break
    if len(r) == len(hdr):
        data.append(r)
isrc, ismp, ia = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ismp]) for r in data)
print('total samples', tot)
agg = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in stall_cols}
print(sorted(agg.items(), key=lambda kv: -kv[1])[:10])
idx = sorted(range(len(data)), key=lambda i: -int(data[i][ismp]))[:N]
for i in sorted(idx):
    r = data[i]
    st = sorted(((hdr[j], int(r[j] or 0)) for j in stall_cols), key=lambda kv: -kv[1])[:3]
    print(f'{i:5d} smp={int(r[ismp]):5d} exec={int(r[ia]):8d}  {r[isrc].strip()[:70]:70s} {st}')
