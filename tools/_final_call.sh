mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_n1.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'dropin',d['e2e']['through_dropin']['value'],'ms/step',d['ms_per_step'])
print('roofline',d['roofline']['achieved'],d['roofline']['frac'],d['roofline']['frac_of_burst_peak'],d['roofline']['traffic'])
print('fixed',{k:d['fixed'].get(k) for k in ['seconds','features_s','finish_ms','frechet_ms']})
print('parity',{k:v for k,v in d['parity'].items() if k not in ('note',)})
print('frechet',d['frechet'],'cpu',d['cpu_baseline']['value'],'clocks',d['clocks'])
PY
