#!/bin/bash
set -x
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python tests/gpu_rate.py bazin,powerlaw 20000 2
timeout 300 python tests/gpu_rate.py bazin,powerlaw 1000000 1
timeout 900 python bench.py --config c1 --full-cpu > gpurun_out/bench_r02_c1.json 2> gpurun_out/bench_r02_c1.err; tail -2 gpurun_out/bench_r02_c1.err
timeout 600 python bench.py --config c2 > gpurun_out/bench_r02_c2.json 2> gpurun_out/bench_r02_c2.err; tail -2 gpurun_out/bench_r02_c2.err
timeout 600 python bench.py --config c3 > gpurun_out/bench_r02_c3.json 2> gpurun_out/bench_r02_c3.err; tail -2 gpurun_out/bench_r02_c3.err
python bench.py --objects 200000 --steps 2 --warmup 1 --no-cpu --gp1d-objects 0 > gpurun_out/plain_launches.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --objects 200000 --steps 2 --warmup 1 --no-cpu --gp1d-objects 0 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
