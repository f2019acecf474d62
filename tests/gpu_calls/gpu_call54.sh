#!/bin/bash
set -x
timeout 900 python bench.py --config c1 --full-cpu > gpurun_out/bench_r02_c1.json 2> gpurun_out/bench_r02_c1.err; tail -1 gpurun_out/bench_r02_c1.err
timeout 600 python bench.py --config c2 > gpurun_out/bench_r02_c2.json 2> gpurun_out/bench_r02_c2.err; tail -1 gpurun_out/bench_r02_c2.err
