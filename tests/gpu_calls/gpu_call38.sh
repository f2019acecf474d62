#!/bin/bash
set -x
MLRN_LIB=variants/lib_gpphase.so timeout 300 python tests/gpu_gp_phases.py 3000
