#!/bin/bash
set -x
timeout 300 python tests/gpu_rate.py gp2d 100000 3
MLRN_LIB=variants/lib_gpnods.so timeout 300 python tests/gpu_rate.py gp2d 100000 3
