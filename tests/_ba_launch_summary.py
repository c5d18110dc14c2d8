import csv, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10 and r[0].isdigit()]
W=sys.argv[2]
idx=[i for i,r in enumerate(rows) if "ba_reset_oob" in r[4] and (", %s," % W) in r[8]]
seg=rows[idx[-1]:]
agg={}
for r in seg:
    name=r[4].split("(")[0].replace("void ","").replace("sdv::",""); agg.setdefault(name,[]).append(int(r[-1]))
tot=sum(sum(v) for v in agg.values()); print("launches",len(seg),"total us",tot/1e3)
for k,v in sorted(agg.items(), key=lambda x:-sum(x[1])): print(f"{k:28s} n={len(v):3d} avg us {sum(v)/len(v)/1e3:8.1f} max {max(v)/1e3:8.1f} total us {sum(v)/1e3:9.1f} share {sum(v)/tot:.3f}")
