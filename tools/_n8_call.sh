mkdir -p gpurun_out
for n in 8 2; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 5 --no-ingest > gpurun_out/bench_r02_n$n.json 2> gpurun_out/bench_r02_n$n.err; echo "n$n rc=$?"
done
python - <<'PY'
import json
for n in (8,2):
    d=json.load(open('gpurun_out/bench_r02_n%d.json'%n))
    print(n, d['value'], d['e2e']['value'], d['finish_ms'], {k:d['fixed'].get(k) for k in ['seconds','features_s','finish_ms','frechet_ms','fid','is_mean']})
PY
