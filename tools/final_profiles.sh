#!/bin/bash
# Regenerates the evidence under profiles/ on a GPU box (outputs land in gpurun_out/; copy the summaries afterwards).
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_r01_n1.err
TISE_PROFILE_DUMP=gpurun_out/layers_r01_current.txt timeout 200 python tools/profile_trunk.py 512 > /dev/null 2>&1
TISE_UMMA_DEBUG=64 timeout 200 python tools/profile_trunk.py 512 2> gpurun_out/roles_r01_current.raw > /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:"conv_umma|conv_flat|stem|pool|resize|moments|logits|is_update" -c 700 --csv \
  --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-frechet > /dev/null 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum \
  --clock-control none -k regex:"conv_umma|conv_flat" -s 132 -c 66 --csv --log-file gpurun_out/ncu_conv_r01.csv \
  python tools/profile_trunk.py 512 > /dev/null 2>&1
echo profiles done
