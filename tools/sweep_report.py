import sys
names=sys.argv[1].split(',')
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.06
def load(n):
    return [(float(ms),float(gf),tag.strip()) for ms,gf,tag in (l.split(None,2) for l in open(f'gpurun_out/sw_{n}.txt'))]
data={n:load(n) for n in names}
base=data[names[0]]
print(f"{'layer':46s}"+"".join(f"{n:>8s}" for n in names))
seen={}; tot={n:0 for n in names}
for i,(ms,gf,tag) in enumerate(base):
    for n in names: tot[n]+=data[n][i][0]
    if tag in seen: continue
    seen[tag]=1
    if ms<thr: continue
    print(f"{tag[:46]:46s}"+"".join(f"{data[n][i][0]*1000:8.0f}" for n in names))
print(f"{'TOTAL ms':46s}"+"".join(f"{tot[n]:8.2f}" for n in names))
