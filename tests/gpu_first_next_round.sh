#!/bin/bash
# First GPU call of the next round: validate and time what round 1 prepared on the CPU.
#   gpurun --timeout 600 -- 'bash tests/gpu_first_next_round.sh > gpurun_out/first.log 2>&1; tail -30 gpurun_out/first.log'
set -x
MLRN_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_research.py -m gpu -q
timeout 60 python tests/gpu_gp1d_rate.py 100000 2
MLRN_GP1D_TILED=1 timeout 60 python tests/gpu_gp1d_rate.py 100000 2
timeout 90 python tests/gpu_research_rate.py 100000 2
