#!/bin/bash
set -x
timeout 300 python tests/gpu_rate.py bazin,powerlaw 1000000 1
timeout 300 python tests/gpu_rate.py bazin,powerlaw 10178 2
timeout 300 python tests/gpu_rate.py gp2d 100000 2
MLRN_LIB=variants/lib_gpilp4.so timeout 300 python tests/gpu_rate.py gp2d 100000 2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_stat_warp' -c 1 -f -o /tmp/prof_e python tests/gpu_rate.py statistical 200000 0 > gpurun_out/ncu_r02e.log 2>&1
ncu -i /tmp/prof_e.ncu-rep --page raw --csv > gpurun_out/prof_r02e_raw.csv 2>/dev/null
ncu -i /tmp/prof_e.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > gpurun_out/prof_r02e_source.csv.gz
