mkdir -p gpurun_out
for cfg in "base" "TISE_UMMA_MSUB1P=128" "TISE_UMMA_MSUB1P=96" "TISE_UMMA_GROUP=1" "TISE_UMMA_GROUP=3" "TISE_UMMA_GROUP=4"; do
  f=gpurun_out/layers_sweep_$(echo $cfg | tr '=' '_').txt
  if [ "$cfg" = "base" ]; then TISE_PROFILE_DUMP=$f timeout 200 python tools/profile_trunk.py 1024 > /dev/null 2>&1
  else env $cfg TISE_PROFILE_DUMP=$f timeout 200 python tools/profile_trunk.py 1024 > /dev/null 2>&1; fi
  echo "$cfg: $(awk '{s+=$1} END {print s}' $f) ms"
done
