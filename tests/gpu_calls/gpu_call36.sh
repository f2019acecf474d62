#!/bin/bash
set -x
timeout 300 python tests/gpu_rate.py gp2d 100000 2
MLRN_LIB=variants/lib_gpskip.so timeout 300 python tests/gpu_rate.py gp2d 100000 2
