#!/bin/bash
set -x
nvidia-smi --query-gpu=index,memory.total --format=csv
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_r02_c5_2gpu.json 2> gpurun_out/bench_r02_c5_2gpu.err; tail -5 gpurun_out/bench_r02_c5_2gpu.err
nvidia-smi --query-gpu=index,memory.used --format=csv
