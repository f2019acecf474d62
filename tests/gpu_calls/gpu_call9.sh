#!/bin/bash
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --objects 100000 --steps 2 --warmup 1 --gp1d-objects 20000 --cpu-objects 64 > gpurun_out/bench_100k.json 2> gpurun_out/bench_100k.err; tail -3 gpurun_out/bench_100k.err
timeout 600 python bench.py --config c1 --objects 2000 --steps 1 --cpu-objects 128 > gpurun_out/bench_c1_small.json 2> gpurun_out/bench_c1_small.err; tail -3 gpurun_out/bench_c1_small.err
