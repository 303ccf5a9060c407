#!/bin/bash
# tools/build_variant.sh <name> <extra nvcc -D flags...>: builds experiments/libps_<name>.so with AOT group 0 (fast) only
set -e
name=$1; shift
cd /root/repo/phasespace_b200/csrc
F="-gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC"
mkdir -p /tmp/t/var/$name
nvcc $F "$@" -c ps_api.cu -o /tmp/t/var/$name/api.o 2>/dev/null
nvcc $F "$@" -DPS_AOT_GROUP=${PS_GROUP:-0} -DPS_EXACT=0 -c ps_aot.cu -o /tmp/t/var/$name/aot.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o /root/repo/experiments/libps_$name.so /tmp/t/var/$name/api.o /tmp/t/var/$name/aot.o -ldl
echo "$name: $(cuobjdump -res-usage /tmp/t/var/$name/aot.o | grep -A1 'Li2ELi0ELin1EJEEEEEELb0ELb0ELi0EEEv8PsParams' | grep -oE 'REG:[0-9]+|STACK:[0-9]+' | tr '\n' ' ')"
