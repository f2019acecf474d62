#!/bin/bash
set -x
timeout 300 python tests/gpu_rate.py gp2d 100000 2
MLRN_LIB=variants/lib_gphelp2.so timeout 300 python tests/gpu_rate.py gp2d 100000 2
MLRN_LIB=variants/lib_gphelp2ph.so timeout 300 python tests/gpu_gp_phases.py 3000 | tail -22
