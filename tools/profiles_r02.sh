#!/bin/bash
# Regenerates the round-2 evidence under profiles/ on a GPU box (outputs land in gpurun_out/; copy the summaries afterwards).
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --workload cfg4 --steps 10 --warmup 3 > gpurun_out/bench_r02_cfg4.json 2> gpurun_out/bench_r02_cfg4.err; echo "cfg4 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err; echo "ref rc=$?"
TISE_PROFILE_DUMP=gpurun_out/layers_r02.txt timeout 200 python tools/profile_trunk.py 1024 > /dev/null 2>&1
# launch list of ONE steady-state step (1024 images) with DRAM / L2 / tensor-pipe counters
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum \
  --clock-control none --csv --log-file gpurun_out/launches_r02.csv python tools/run_step.py --batch 1024 --steps 1 > gpurun_out/launches_r02.out 2>&1
python tools/ncu_step_summary.py gpurun_out/launches_r02.csv > gpurun_out/launches_r02_summary.txt
grep -aE '^"ID"|conv_umma|conv_flat' gpurun_out/launches_r02.csv > gpurun_out/ncu_conv_r02.csv
python tools/ncu_conv_summary.py gpurun_out/ncu_conv_r02.csv gpurun_out/ncu_conv_r02.json > /dev/null
# one full capture of the largest tcgen05 convolution (Conv2d_4a, second conv_umma launch of the step) and of a 17 x 17 1x7 layer
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_umma --launch-skip 1 --launch-count 1 \
  -o gpurun_out/conv4a_r02 -f python tools/run_step.py --batch 1024 --steps 1 > gpurun_out/conv4a_r02.out 2>&1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_umma --launch-skip 38 --launch-count 1 \
  -o gpurun_out/conv17_r02 -f python tools/run_step.py --batch 1024 --steps 1 > gpurun_out/conv17_r02.out 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tri_resident --launch-count 1 -o gpurun_out/tri_r02 -f python tools/frechet_run.py --reps 1 > gpurun_out/tri_r02.out 2>&1
ls -la gpurun_out/*.ncu-rep
echo profiles done
