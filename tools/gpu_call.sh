#!/bin/bash
# One gpurun call: the whole -m gpu suite, then bench (cfg2).  Logs land in gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests rc=$?"; grep -aoE "cfg[12][^|]*\|[^\n]*" gpurun_out/gpu_tests.log; tail -4 gpurun_out/gpu_tests.log
timeout 600 python bench.py > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err
echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r02_b.err; head -c 400 gpurun_out/bench_r02_b.json
timeout 300 python bench.py --strict --no-cpu-baseline --no-frechet --no-ingest --fixed 0 > gpurun_out/bench_r02_b_strict.json 2> gpurun_out/bench_r02_b_strict.err
echo "bench strict rc=$?"; head -c 300 gpurun_out/bench_r02_b_strict.json
