#!/bin/bash
set -x
timeout 300 python tests/gpu_rate.py bazin,powerlaw 1000000 1
MLRN_LIB=variants/lib_fit256x2.so timeout 300 python tests/gpu_rate.py bazin,powerlaw 1000000 1
MLRN_LIB=variants/lib_fit128x4.so timeout 300 python tests/gpu_rate.py bazin,powerlaw 1000000 1
