#!/bin/bash
# The C5 batch plan on two GPUs: a 1.6 M-object set, 2 batches of 400 k objects per rank, all-gather per step.
set -x
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --objects 1600000 --batch-objects 400000 --steps 2 --warmup 1 --no-cpu --gp1d-objects 0 > gpurun_out/bench_c5_plan_2gpu.json 2> gpurun_out/bench_c5_plan_2gpu.err; echo rc=$?; tail -3 gpurun_out/bench_c5_plan_2gpu.err; cut -c1-1800 gpurun_out/bench_c5_plan_2gpu.json
