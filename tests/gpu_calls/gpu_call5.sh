#!/bin/bash
set -x
MLRN_LIB=$PWD/variants/lib_r01.so timeout 200 python tests/gpu_rate.py gp2d 100000 2
for v in r01phase gpphase gpphasechol; do
MLRN_LIB=$PWD/variants/lib_$v.so timeout 200 python tests/gpu_gp_phases.py 3000
done
