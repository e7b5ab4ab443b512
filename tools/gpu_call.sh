#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests rc=$?"; grep -aoE "cfg[12] [^|]*\|[^\n]*" gpurun_out/gpu_tests.log; tail -4 gpurun_out/gpu_tests.log
timeout 600 python bench.py > gpurun_out/bench_r02_d.json 2> gpurun_out/bench_r02_d.err
echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r02_d.err; head -c 300 gpurun_out/bench_r02_d.json
TISE_PROFILE_DUMP=gpurun_out/layers_r02_d.txt timeout 200 python tools/profile_trunk.py 1024 > /dev/null 2>&1
timeout 300 python tools/cli_timing.py --n 8192 > gpurun_out/cli_timing_r02.txt 2>&1; tail -3 gpurun_out/cli_timing_r02.txt
