#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_gp2d.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tests/gpu_rate.py gp2d 100000 2
MLRN_LIB=variants/lib_gpnoswz.so timeout 300 python tests/gpu_rate.py gp2d 100000 2
