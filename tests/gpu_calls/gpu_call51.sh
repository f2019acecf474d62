#!/bin/bash
set -x
MLRN_LIB=variants/lib_gp640.so timeout 300 python tests/gpu_rate.py gp2d 100000 2
MLRN_LIB=variants/lib_gp384.so timeout 300 python tests/gpu_rate.py gp2d 100000 2
