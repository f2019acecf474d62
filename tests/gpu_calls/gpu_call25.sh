#!/bin/bash
# fits: products pinned against contraction (numpy semantics), power law as exp(k ln x): parity classes + rates
set -x
timeout 900 python -m pytest tests/test_gpu_fits.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python tests/gpu_fit_report.py 1000 2>&1 | head -12
timeout 300 python tests/gpu_rate.py bazin,powerlaw 1000000 1
