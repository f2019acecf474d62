#!/bin/bash
# ncu of the final k_gp2d (after the bank-conflict swizzle and the padding skip) and of the final Bazin stage 0 (prologue, new caps)
set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gp2d' -c 1 -f -o /tmp/prof_h python tests/gpu_rate.py gp2d 20000 0 > gpurun_out/ncu_r02h.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_fit_stage' -c 1 -f -o /tmp/prof_i python tests/gpu_rate.py bazin 200000 0 > gpurun_out/ncu_r02i.log 2>&1
for x in h i; do
  ncu -i /tmp/prof_$x.ncu-rep --page raw --csv > gpurun_out/prof_r02${x}_raw.csv 2>/dev/null
done
