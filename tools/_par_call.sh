mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_baseline.py tests/test_gpu_trunk.py -q -m gpu -s 2>&1 | grep -aE "cfg[12] |passed|failed" | sed 's/^[.s]*//'
timeout 200 python tools/diag_policy.py 26 200 2>&1 | grep -E "crops|fid  "
