import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
n=int(sys.argv[2]) if len(sys.argv)>2 else 25
ctx=int(sys.argv[3]) if len(sys.argv)>3 else 0
hdr=rows[1]; ci={x:i for i,x in enumerate(hdr)}
data=rows[2:]
tot=sum(float(r[ci['# Samples']] or 0) for r in data)
print("total samples",tot)
stallcols=[x for x in hdr if x.startswith('stall_') and 'Not Issued' not in x]
top=sorted(enumerate(data), key=lambda t:-float(t[1][ci['# Samples']] or 0))[:n]
show=set()
for i,_ in top:
    for k in range(i-ctx,i+1): show.add(k)
for i in sorted(show):
    r=data[i]
    reasons={x:float(r[ci[x]] or 0) for x in stallcols if r[ci[x]] not in ('0','','-')}
    rs=sorted(reasons.items(), key=lambda kv:-kv[1])[:2]
    print(i, f"{r[ci['# Samples']]:>6s} {r[ci['Instructions Executed']]:>8s} | {r[ci['Source']].strip()[:70]:70s} {[(k[6:],int(v)) for k,v in rs]}")
