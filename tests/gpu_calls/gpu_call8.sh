#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_fits.py -m gpu -q -x
timeout 200 python tests/gpu_rate.py bazin,powerlaw 20000 2
timeout 200 python tests/gpu_rate.py bazin,powerlaw 200000 2
timeout 300 python tests/gpu_rate.py bazin,powerlaw 1000000 1
for v in fitcapsA fitcapsB; do
MLRN_LIB=$PWD/variants/lib_$v.so timeout 200 python tests/gpu_rate.py bazin 20000 2
MLRN_LIB=$PWD/variants/lib_$v.so timeout 200 python tests/gpu_rate.py bazin 200000 2
done
MLRN_LIB=$PWD/variants/lib_r01.so timeout 300 python tests/gpu_rate.py bazin,powerlaw 1000000 1
