mkdir -p gpurun_out
python tools/run_step.py --batch 1024 --steps 4 2>&1 | tail -1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/launches_r02.csv python tools/run_step.py --batch 1024 --steps 1 > gpurun_out/launches_r02.out 2>&1
python tools/ncu_step_summary.py gpurun_out/launches_r02.csv | tee gpurun_out/launches_r02_summary.txt
