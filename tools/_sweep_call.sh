TISE_PROFILE_DUMP=gpurun_out/layers_g4.txt timeout 200 python tools/profile_trunk.py 1024 > /dev/null 2>&1
TISE_UMMA_G4=0 TISE_PROFILE_DUMP=gpurun_out/layers_g4off.txt timeout 200 python tools/profile_trunk.py 1024 > /dev/null 2>&1
echo "G4 total $(awk '{s+=$1} END {print s}' gpurun_out/layers_g4.txt)  off $(awk '{s+=$1} END {print s}' gpurun_out/layers_g4off.txt)"
timeout 500 python -m pytest tests/test_gpu_trunk.py tests/test_gpu_kernels.py tests/test_gpu_e2e.py -q -m gpu 2>&1 | tail -2
