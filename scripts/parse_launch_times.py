import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; iv=hdr.index("Metric Value"); ik=hdr.index("Kernel Name"); ig=hdr.index("Grid Size"); ib=hdr.index("Block Size")
d=collections.defaultdict(list)
for r in rows[1:]:
    d[(r[ik][:50], r[ig], r[ib])].append(float(r[iv].replace(",","")))
for k,v in d.items(): print(k, len(v), "median %.1f us min %.1f max %.1f" % (sorted(v)[len(v)//2]/1e3, min(v)/1e3, max(v)/1e3))
