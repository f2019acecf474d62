#!/bin/bash
set -x
cd gpurun_out && rm -f prof_*.ncu-rep && cd ..
python tests/gpu_rate.py statistical,bazin,gp2d 20000 0 > gpurun_out/plain_r02.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k regex:'k_stat|k_fit_stage' -c 6 -f -o /tmp/prof_a python tests/gpu_rate.py statistical,bazin 20000 0 > gpurun_out/ncu_r02a.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_gp2d' -c 1 -f -o /tmp/prof_b python tests/gpu_rate.py gp2d 20000 0 > gpurun_out/ncu_r02b.log 2>&1
for x in a b; do
  ncu -i /tmp/prof_$x.ncu-rep --page raw --csv > gpurun_out/prof_r02${x}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$x.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > gpurun_out/prof_r02${x}_source.csv.gz
done
ls -la gpurun_out/ /tmp/prof_*.ncu-rep
