"""per-source-line stall samples of an `ncu --set full --import-source on` report (development aid):
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv ; python tests/ncu_hotlines.py x.csv [N]"""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if len(r)>5 and r[0]=='Line No'][0]
h=rows[hi]
ss=h.index('# Samples'); ie=h.index('Instructions Executed')
data=[]
for r in rows[hi+1:]:
    if len(r)<len(h) or r[0]=='' : continue
    try: data.append((int(r[ss]), int(r[ie]), int(r[0]), r[1].strip()[:110], r))
    except: pass
tot=sum(d[0] for d in data); toti=sum(d[1] for d in data)
print("total samples",tot,"inst",toti)
stall_cols=[i for i,n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
agg={}
for d in data:
    for i in stall_cols:
        v=int(d[4][i]) if d[4][i].isdigit() else 0
        agg[h[i][6:]]=agg.get(h[i][6:],0)+v
print(sorted(agg.items(), key=lambda x:-x[1])[:8])
for d in sorted(data,reverse=True)[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
    st=sorted([(int(d[4][i]) if d[4][i].isdigit() else 0, h[i][6:]) for i in stall_cols], reverse=True)[:3]
    print("%5.1f%% smp %5.1f%% inst L%-4d %-100s %s"%(100*d[0]/tot,100*d[1]/toti,d[2],d[3], " ".join("%s:%d"%(n,c) for c,n in st)))
