#!/bin/bash
# plain run, then one `ncu --set full` capture of every generator kernel of tools/profile_configs.py
set -e
mkdir -p gpurun_out
python tools/profile_configs.py 1e7 > gpurun_out/profile_full_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:genbod -o gpurun_out/prof_configs -f python tools/profile_configs.py 1e7 > gpurun_out/profile_full_ncu.log 2>&1
tail -3 gpurun_out/profile_full_ncu.log
ls -la gpurun_out/prof_configs.ncu-rep
