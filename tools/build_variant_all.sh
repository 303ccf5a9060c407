#!/bin/bash
# tools/build_variant_all.sh <name> <extra nvcc -D flags...>: builds experiments/libps_<name>.so with EVERY ahead-of-time
# group of the fast variant (no exact kernels), for A/B timing with PS_B200_LIB=experiments/libps_<name>.so
set -e
name=$1; shift
cd /root/repo/phasespace_b200/csrc
F="-gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC"
d=/tmp/t/var/$name; mkdir -p $d /root/repo/experiments
ngroups=$(grep -oE 'PS_AOT_NUM_GROUPS [0-9]+' ps_aot_trees.inc | grep -oE '[0-9]+')
nvcc $F "$@" -c ps_api.cu -o $d/api.o 2>/dev/null &
for g in $(seq 0 $((ngroups-1))); do
  nvcc $F "$@" -DPS_AOT_GROUP=$g -DPS_EXACT=0 -c ps_aot.cu -o $d/aot$g.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o /root/repo/experiments/libps_$name.so $d/api.o $d/aot*.o -ldl
echo "built experiments/libps_$name.so"
