python tools/frechet_run.py --reps 3 2>&1 | tail -2
python tools/frechet_run.py --d 768 --n 2000 --reps 2 2>&1 | tail -1
python tools/frechet_run.py --d 192 --n 500 --reps 2 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -m gpu -k "frechet or fid or stress" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/frechet_launches_r02.csv python tools/frechet_run.py --reps 1 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/frechet_launches_r02.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    n=r[ki].split('(')[0]
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(v[1] for v in agg.values())
for n,(c,tt) in sorted(agg.items(), key=lambda x:-x[1][1]): print('%-44s %5d %12.1f us %6.3f'%(n[:44],c,tt/1e3,tt/tot))
print('total_us %.1f'%(tot/1e3))
PY
