#!/bin/bash
for m in statistical colors shape tde physics research bazin powerlaw gp2d gp1d; do
  for k in nan pinf ninf; do
    timeout 60 python tests/gpu_nonfinite_probe.py $m $k 2>&1 | tail -1 | cut -c1-200
  done
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
