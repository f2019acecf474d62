#!/bin/bash
# K2 (exact mean/sd, hoisted loads) parity + rate, register cap variant; K4 work order and register-resident factor
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fits.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -6
timeout 300 python tests/gpu_rate.py statistical 1000000 3
MLRN_LIB=variants/lib_sw5.so timeout 300 python tests/gpu_rate.py statistical 1000000 3
for n in 10178 200000 1000000; do
  timeout 300 python tests/gpu_rate.py bazin,powerlaw $n 2
  MLRN_FIT_UNORDERED=1 timeout 300 python tests/gpu_rate.py bazin,powerlaw $n 2
  MLRN_LIB=variants/lib_regf.so timeout 300 python tests/gpu_rate.py bazin,powerlaw $n 2
done
