#!/bin/bash
# One gpurun call: parity at BASELINE sizes, pass-policy probe, bench (cfg2 + cfg4).  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/smi.txt
timeout 900 python -m pytest tests/test_gpu_parity_baseline.py tests/test_gpu_e2e.py -x -q -m gpu -s > gpurun_out/parity_tests.log 2>&1
echo "parity tests rc=$?" | tee -a gpurun_out/parity_tests.log
timeout 600 python tools/pass_policy_probe.py 1000 gpurun_out/pass_policy.json > gpurun_out/pass_policy.log 2>&1
echo "probe rc=$?"
timeout 600 python bench.py > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r02_a.err
timeout 600 python bench.py --workload cfg4 --steps 10 > gpurun_out/bench_r02_cfg4_a.json 2> gpurun_out/bench_r02_cfg4_a.err
echo "cfg4 rc=$?"; tail -c 600 gpurun_out/bench_r02_cfg4_a.err
tail -5 gpurun_out/parity_tests.log
