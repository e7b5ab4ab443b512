"""Summarises a TISE_PROFILE_DUMP file: python tools/layers.py file [n_microbatches]"""
import sys
from collections import OrderedDict
rows = [l.split(None, 2) for l in open(sys.argv[1])]
nmb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
agg = OrderedDict()
for ms, gf, tag in rows:
    a = agg.setdefault(tag.strip(), [0, 0.0, 0.0]); a[0] += 1; a[1] += float(ms); a[2] += float(gf)
tot = sum(a[1] for a in agg.values())
print("total ms per microbatch %.3f" % (tot / nmb))
for tag, (n, ms, gf) in agg.items():
    if ms / tot < float(sys.argv[3]) if len(sys.argv) > 3 else False: continue
    print(f"{ms/n*1000:8.1f} us {gf/ms if ms else 0:7.1f} TF/s share {ms/tot:5.3f} x{n//nmb:2d} {tag}")
