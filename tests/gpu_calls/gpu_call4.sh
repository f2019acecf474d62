#!/bin/bash
set -x
for v in gpphase gpphase0 gpphasebase; do
MLRN_LIB=$PWD/variants/lib_$v.so timeout 200 python tests/gpu_gp_phases.py 3000
done
