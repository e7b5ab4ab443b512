#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests rc=$?"; grep -aoE "cfg[12] [^|]*\|[^\n]*" gpurun_out/gpu_tests.log; tail -4 gpurun_out/gpu_tests.log
timeout 600 python tools/pass_policy_probe.py 1000 gpurun_out/pass_policy.json > gpurun_out/pass_policy.log 2>&1
echo "probe rc=$?"; grep policy gpurun_out/pass_policy.log
timeout 600 python bench.py > gpurun_out/bench_r02_c.json 2> gpurun_out/bench_r02_c.err
echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r02_c.err; head -c 400 gpurun_out/bench_r02_c.json
