#!/bin/bash
# plain run first, then the ncu metrics pass over every generator kernel (one launch each)
set -e
mkdir -p gpurun_out
python tools/profile_configs.py > gpurun_out/profile_configs_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,dram__bytes_write.sum,dram__bytes_read.sum,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio \
    --clock-control none -k regex:genbod --csv --log-file gpurun_out/profile_configs.csv python tools/profile_configs.py > gpurun_out/profile_configs_ncu.log 2>&1
tail -3 gpurun_out/profile_configs_ncu.log
