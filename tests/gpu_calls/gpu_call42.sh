#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "non_finite or corner" 2>&1 | tail -12
