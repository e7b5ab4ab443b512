TISE_FRECHET_DEBUG=64 timeout 60 python tools/frechet_run.py --reps 2 2>&1 | grep -E "tri_resident|rep 1" | tail -2
python tools/frechet_run.py --reps 3 2>&1 | tail -2
TISE_FRECHET_RESIDENT=2 timeout 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu -k "frechet" 2>&1 | tail -2
