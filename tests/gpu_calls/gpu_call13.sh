#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_interactions.py tests/test_gpu_frames.py -m gpu -x -q 2>&1 | tail -5
timeout 1500 python bench.py > gpurun_out/bench_r02_1m.json 2> gpurun_out/bench_r02_1m.err; tail -3 gpurun_out/bench_r02_1m.err
