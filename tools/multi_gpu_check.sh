#!/bin/bash
# N-GPU check (gpurun --gpus N): the headline bench under torchrun, then config 5 (1e10 candidates, unweighting)
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu_v5.log 2>&1
tail -1 gpurun_out/bench_${N}gpu_v5.log | cut -c1-1500
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/bench_config5.py > gpurun_out/config5_${N}gpu_v5.log 2>&1
tail -1 gpurun_out/config5_${N}gpu_v5.log
