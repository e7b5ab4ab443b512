mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err; echo "n2 rc=$?"
tail -c 600 gpurun_out/bench_r02_n2.err
timeout 600 python -m pytest tests/test_gpu_multirank_cli.py -q -m gpu 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r02_reference_n2.json 2>/dev/null; echo "ref n2 rc=$?"; head -c 300 gpurun_out/bench_r02_reference_n2.json
