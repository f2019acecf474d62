#!/bin/bash
# final single-GPU run of round 2: full GPU suite, smoke, default bench line, the three small configs
set -x
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 1500 python bench.py > gpurun_out/bench_r02_final_1m.json 2> gpurun_out/bench_r02_final_1m.err; tail -2 gpurun_out/bench_r02_final_1m.err
timeout 900 python bench.py --config c1 --full-cpu > gpurun_out/bench_r02_c1.json 2> gpurun_out/bench_r02_c1.err; tail -2 gpurun_out/bench_r02_c1.err
timeout 600 python bench.py --config c2 > gpurun_out/bench_r02_c2.json 2> gpurun_out/bench_r02_c2.err; tail -2 gpurun_out/bench_r02_c2.err
timeout 600 python bench.py --config c3 > gpurun_out/bench_r02_c3.json 2> gpurun_out/bench_r02_c3.err; tail -2 gpurun_out/bench_r02_c3.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err; tail -2 gpurun_out/bench_r02_reference.err
