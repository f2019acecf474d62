#!/bin/bash
# The C5 batch plan of bench.py on one GPU (4 batches of 500 k objects of a 2 M-object set), the C4 line with every leg at
# 200 k objects, then the GPU tests and smoke on the committed build.
set -x
timeout 300 python bench.py --config c5 --objects 2000000 --batch-objects 500000 --steps 4 --warmup 1 --no-cpu --gp1d-objects 0 > gpurun_out/bench_c5_plan_1gpu.json 2> gpurun_out/bench_c5_plan_1gpu.err; echo rc=$?; tail -2 gpurun_out/bench_c5_plan_1gpu.err; cut -c1-1500 gpurun_out/bench_c5_plan_1gpu.json
timeout 300 python bench.py --objects 200000 --steps 2 --warmup 1 > gpurun_out/bench_c4_200k.json 2> gpurun_out/bench_c4_200k.err; echo rc=$?; tail -2 gpurun_out/bench_c4_200k.err; cut -c1-600 gpurun_out/bench_c4_200k.json
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-300
