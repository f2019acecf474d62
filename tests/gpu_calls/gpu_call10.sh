#!/bin/bash
set -x
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python tests/gpu_rate.py statistical,colors,shape,tde,physics,research,gp1d 100000 2
timeout 600 python bench.py --objects 200000 --steps 2 --warmup 1 --gp1d-objects 0 --no-cpu > gpurun_out/bench_200k.json 2> gpurun_out/bench_200k.err; tail -3 gpurun_out/bench_200k.err
