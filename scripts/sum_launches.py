import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
hdr=rows[hi]; ix={h:i for i,h in enumerate(hdr)}
agg=collections.OrderedDict(); tot=0
for r in rows[hi+1:]:
    if len(r)!=len(hdr): continue
    v=float(r[ix["Metric Value"]]); unit=r[ix["Metric Unit"]]
    us = v/1000.0 if unit in ("nsecond","ns") else (v if unit in ("usecond","us") else v*1000.0)
    a=agg.setdefault(r[ix["Kernel Name"]][:100],[0,0.0]); a[0]+=1; a[1]+=us; tot+=us
print("total us", tot, "launches", sum(a[0] for a in agg.values()))
for k,(n,us) in sorted(agg.items(), key=lambda x:-x[1][1])[:28]: print(f"{n:6d} {us:10.1f} {100*us/tot:5.1f}%  {k}")
