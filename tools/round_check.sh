#!/bin/bash
# Round check on one B200: GPU tests, the bench line, the fp64 pipe microbenchmark, then the two ncu passes over the
# same short bench command (launch list, and one --set full capture of the generator kernel).
#   tools/round_check.sh [tag]     (tag names the output files, default r02)
TAG=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_$TAG.log 2>gpurun_out/bench_$TAG.err; tail -c 3000 gpurun_out/bench_$TAG.log
[ -x tools/bin/fp64_bench ] && ./tools/bin/fp64_bench > gpurun_out/fp64_bench.log 2>&1
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-sustained"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:genbod -s 3 -c 1 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu2_$TAG.log 2>&1
tail -2 gpurun_out/ncu2_$TAG.log
