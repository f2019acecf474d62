#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_fits.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tests/gpu_rate.py bazin,powerlaw 10178 2
timeout 300 python tests/gpu_rate.py bazin,powerlaw 20000 2
timeout 300 python tests/gpu_rate.py bazin,powerlaw 200000 2
timeout 300 python tests/gpu_rate.py bazin,powerlaw 1000000 1
