#!/bin/bash
set -x
timeout 600 python tests/gpu_overlap.py 200000
MLRN_LIB=variants/lib_fit128.so timeout 600 python tests/gpu_overlap.py 200000
MLRN_LIB=variants/lib_fit128.so timeout 300 python tests/gpu_rate.py bazin,powerlaw 200000 1
