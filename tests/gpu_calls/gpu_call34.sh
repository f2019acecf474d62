#!/bin/bash
# two GPUs: the NCCL test of the sharded entry and the C5 bench line (one fixed 10^7-object set over 2 ranks)
set -x
timeout 600 python -m pytest tests/test_shard.py -m gpu -x -q 2>&1 | tail -3
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_r02_c5_2gpu.json 2> gpurun_out/bench_r02_c5_2gpu.err; tail -3 gpurun_out/bench_r02_c5_2gpu.err
