#!/bin/bash
set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_shard.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --objects 400000 --no-cpu > gpurun_out/bench_2gpu_400k.json 2> gpurun_out/bench_2gpu_400k.err; tail -5 gpurun_out/bench_2gpu_400k.err
