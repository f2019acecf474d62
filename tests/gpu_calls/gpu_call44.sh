#!/bin/bash
for k in t_nan t_inf e_nan e_inf e_zero e_neg; do timeout 120 python tests/gpu_nonfinite_probe2.py $k 2>&1 | tail -11 | cut -c1-220; done
