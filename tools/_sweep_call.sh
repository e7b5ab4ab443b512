TISE_PROFILE_DUMP=gpurun_out/layers_stem4.txt timeout 200 python tools/profile_trunk.py 1024 > /dev/null 2>&1
head -1 gpurun_out/layers_stem4.txt; echo "total $(awk '{s+=$1} END {print s}' gpurun_out/layers_stem4.txt)"
timeout 400 python -m pytest tests/test_gpu_trunk.py tests/test_gpu_kernels.py -q -m gpu 2>&1 | tail -2
