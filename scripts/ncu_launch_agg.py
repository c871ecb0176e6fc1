import csv, sys
from collections import OrderedDict
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
skip=int(sys.argv[2]) if len(sys.argv)>2 else 0
agg=OrderedDict()
data=rows[1:]
data=data[int(len(data)*skip/100):]
for r in data:
    v=float(r[vi].replace(',','')); u=r[ui]
    v = v/1e6 if u=='ns' else (v/1e3 if u=='us' else v)
    k=r[ki].split('(')[0][-70:]
    agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
for k,(c,v) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:18]:
    print('%-72s %4d  %9.3f ms  %5.1f%%' % (k,c,v,100*v/tot))
print('total', tot)
