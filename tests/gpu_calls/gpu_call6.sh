#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_gp2d.py -m gpu -q -x
timeout 200 python tests/gpu_rate.py gp2d 100000 2
for v in gpshare r01; do
MLRN_LIB=$PWD/variants/lib_$v.so timeout 200 python tests/gpu_rate.py gp2d 100000 2
done
MLRN_LIB=$PWD/variants/lib_gpphase.so timeout 200 python tests/gpu_gp_phases.py 3000
