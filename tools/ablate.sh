#!/bin/bash
# timing ablations of conv_umma_kernel (results are garbage for debug != 0; only the timings matter)
mkdir -p gpurun_out
for d in ${ABL:-0 3 4 8 12 7}; do
  TISE_UMMA_DEBUG=$d TISE_PROFILE_DUMP=gpurun_out/abl_$d.txt timeout 300 python tools/profile_trunk.py 256 > /dev/null 2>&1 || echo "debug $d failed"
done
