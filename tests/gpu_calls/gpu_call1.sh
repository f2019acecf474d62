#!/bin/bash
# round 2, GPU call 1: validate what round 1 prepared, and sweep the launch shape of the fit kernels
set -x
MLRN_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_research.py -m gpu -q
timeout 120 python tests/gpu_gp1d_rate.py 100000 2
MLRN_GP1D_TILED=1 timeout 120 python tests/gpu_gp1d_rate.py 100000 2
timeout 120 python tests/gpu_research_rate.py 100000 2
timeout 200 python tests/gpu_rate.py bazin,powerlaw 200000 2
for v in fit256x2 fit128x4 fit256x1 fit128x4nb fit64x8; do
  MLRN_LIB=$PWD/variants/lib_$v.so timeout 200 python tests/gpu_rate.py bazin,powerlaw 200000 2
done
timeout 100 python tests/gpu_rate.py bazin,powerlaw 20000 2
MLRN_LIB=$PWD/variants/lib_fit128x4.so timeout 100 python tests/gpu_rate.py bazin,powerlaw 20000 2
MLRN_LIB=$PWD/variants/lib_gpphase.so timeout 200 python tests/gpu_gp_phases.py 3000
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv
