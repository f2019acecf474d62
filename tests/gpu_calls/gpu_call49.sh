#!/bin/bash
set -x
for v in capsA capsB capsC capsD; do
  MLRN_LIB=variants/lib_$v.so timeout 300 python tests/gpu_rate.py bazin 1000000 1
  MLRN_LIB=variants/lib_$v.so timeout 300 python tests/gpu_rate.py bazin 10178 2
done
