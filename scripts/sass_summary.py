"""Summarise an `ncu --page source --print-source sass --csv` export: per-kernel opcode mix and hot regions."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 500
kern, hdr, data = None, None, []


def flush():
    if not data:
        return
    ia, isrc, ismp = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
    tot = sum(int(r[ia]) for r in data)
    print('==', kern[:90], '| executed', tot, '| sass instrs', len(data))
    ops = collections.Counter()
    for r in data:
        t = r[isrc].split()
        op = t[1] if t[0].startswith('@') else t[0]
        ops[op.split('.')[0]] += int(r[ia])
    print('  ', [(k, f'{100*v/tot:.1f}%') for k, v in ops.most_common(22)])
    for b in range(0, len(data), B):
        s = sum(int(r[ia]) for r in data[b:b + B])
        sm = sum(int(r[ismp]) for r in data[b:b + B])
        if s * 200 > tot:
            print(f'   [{b:6d}] exec {100*s/tot:5.1f}%  samples {sm}')


for r in rows:
    if r and r[0] == 'Kernel Name':
        flush()
        kern, hdr, data = r[1], None, []
    elif r and r[0] == 'Address':
        hdr = r
    elif hdr and len(r) == len(hdr):
        data.append(r)
flush()
