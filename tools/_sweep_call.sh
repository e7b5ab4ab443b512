timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_crop.py -q -m gpu -k "resize or crop" 2>&1 | tail -2
python tools/run_step.py --batch 4096 --steps 4 2>&1 | tail -1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -k regex:resize_fused --csv python tools/run_step.py --batch 1024 --steps 1 2>&1 | grep resize_fused | awk -F'","' '{print $NF}'
