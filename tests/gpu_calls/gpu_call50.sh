#!/bin/bash
set -x
timeout 300 python tests/gpu_rate.py bazin,powerlaw 1000000 1
for v in dcl1 dcl2 dcl3; do
  MLRN_LIB=variants/lib_$v.so timeout 300 python tests/gpu_rate.py powerlaw 1000000 1
  MLRN_LIB=variants/lib_$v.so timeout 300 python tests/gpu_rate.py powerlaw 10178 2
done
for v in coop2 coop8; do
  MLRN_LIB=variants/lib_$v.so timeout 300 python tests/gpu_rate.py bazin 1000000 1
  MLRN_LIB=variants/lib_$v.so timeout 300 python tests/gpu_rate.py bazin 10178 2
done
