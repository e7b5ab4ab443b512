import sys
cols=[int(x) for x in (sys.argv[1].split(',') if len(sys.argv)>1 else "0,3,4,8,12,7".split(','))]
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.07
def load(d):
    return [(float(ms),float(gf),tag.strip()) for ms,gf,tag in (l.split(None,2) for l in open(f'gpurun_out/abl_{d}.txt'))]
data={d:load(d) for d in cols}
base=data[cols[0]]
print("debug bits: 1=noA 2=noB 4=noMMA 8=noEpilogue")
print(f"{'layer':48s}"+"".join(f"{'d'+str(d):>8s}" for d in cols))
seen=set(); tot={d:0 for d in cols}
for i,(ms,gf,tag) in enumerate(base):
    for d in cols: tot[d]+=data[d][i][0]
    if not tag.startswith('conv') or tag in seen or ms<thr: continue
    seen.add(tag)
    print(f"{tag[5:]:48s}"+"".join(f"{data[d][i][0]*1000:8.0f}" for d in cols))
print(f"{'TOTAL ms':48s}"+"".join(f"{tot[d]:8.2f}" for d in cols))
