#!/bin/bash
# fit prologue pass (median + first cost computed up front): parity, bit-identity, rates with and without
set -x
timeout 900 python -m pytest tests/test_gpu_fits.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -4
for n in 10178 1000000; do
  timeout 300 python tests/gpu_rate.py bazin,powerlaw $n 2
  MLRN_FIT_NO_PROLOGUE=1 timeout 300 python tests/gpu_rate.py bazin,powerlaw $n 2
done
