#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "corner or statistical" 2>&1 | tail -5
for v in "" variants/lib_gppo.so variants/lib_gpeh.so variants/lib_gppoeh.so; do
  env ${v:+MLRN_LIB=$v} timeout 300 python tests/gpu_rate.py gp2d 100000 2
done
