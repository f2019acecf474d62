#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gp2d.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tests/gpu_rate.py statistical 1000000 3
MLRN_LIB=variants/lib_sw6.so timeout 300 python tests/gpu_rate.py statistical 1000000 3
