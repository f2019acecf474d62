#!/bin/bash
# GP: evaluation at the optimum not repeated; fits: per-pass reciprocals (div_r)
set -x
timeout 900 python -m pytest tests/test_gpu_fits.py tests/test_gpu_gp2d.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tests/gpu_rate.py gp2d 100000 2
timeout 300 python tests/gpu_rate.py bazin,powerlaw 1000000 1
timeout 300 python tests/gpu_rate.py bazin,powerlaw 10178 2
