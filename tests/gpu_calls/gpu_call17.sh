#!/bin/bash
# K2 rewrite (stat_warp.cuh): parity, rate, ncu; K4 stage-0 ncu at scale (lane utilisation question)
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_frames.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python tests/gpu_rate.py statistical 200000 3
timeout 300 python tests/gpu_rate.py statistical 1000000 3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_stat' -c 2 -f -o /tmp/prof_c python tests/gpu_rate.py statistical 200000 0 > gpurun_out/ncu_r02c.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_fit_stage' -c 1 -f -o /tmp/prof_d python tests/gpu_rate.py bazin 200000 0 > gpurun_out/ncu_r02d.log 2>&1
for x in c d; do
  ncu -i /tmp/prof_$x.ncu-rep --page raw --csv > gpurun_out/prof_r02${x}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$x.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > gpurun_out/prof_r02${x}_source.csv.gz
done
ls -la gpurun_out/ | tail -8
