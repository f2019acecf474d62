#!/bin/bash
# final state of round 2: full GPU suite, smoke, default bench line
set -x
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python bench.py > gpurun_out/bench_r02_final_1m.json 2> gpurun_out/bench_r02_final_1m.err; tail -2 gpurun_out/bench_r02_final_1m.err
timeout 600 python bench.py --config c3 > gpurun_out/bench_r02_c3.json 2> gpurun_out/bench_r02_c3.err; tail -2 gpurun_out/bench_r02_c3.err
