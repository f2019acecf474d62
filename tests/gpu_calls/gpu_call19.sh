#!/bin/bash
# K4: lane lending (idle lanes evaluate the odd rows of a partner's Jacobian pass) vs the same build without it
set -x
timeout 900 python -m pytest tests/test_gpu_fits.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -6
for n in 10178 200000 1000000; do
  timeout 300 python tests/gpu_rate.py bazin,powerlaw $n 2
  MLRN_LIB=variants/lib_nolend.so timeout 300 python tests/gpu_rate.py bazin,powerlaw $n 2
done
