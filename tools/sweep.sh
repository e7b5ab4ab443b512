#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" TISE_PROFILE_DUMP=gpurun_out/sw_$name.txt timeout 300 python tools/profile_trunk.py 256 > gpurun_out/sw_$name.log 2>&1 || { echo "$name FAILED"; tail -3 gpurun_out/sw_$name.log; }; }
run c1 TISE_UMMA_CLUSTER=1
run d15 TISE_UMMA_CLUSTER=1 TISE_UMMA_DEBUG=15
run g1 TISE_UMMA_CLUSTER=1 TISE_UMMA_GROUP=1
run g2 TISE_UMMA_CLUSTER=1 TISE_UMMA_GROUP=2
run c2 TISE_UMMA_CLUSTER=2
