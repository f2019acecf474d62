#!/bin/bash
# final ncu captures of the round-2 kernels + launch list of the bench step
set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stat_warp|k_fit_stage' -c 2 -f -o /tmp/prof_f python tests/gpu_rate.py statistical,bazin 200000 0 > gpurun_out/ncu_r02f.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gp2d' -c 1 -f -o /tmp/prof_g python tests/gpu_rate.py gp2d 20000 0 > gpurun_out/ncu_r02g.log 2>&1
for x in f g; do
  ncu -i /tmp/prof_$x.ncu-rep --page raw --csv > gpurun_out/prof_r02${x}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$x.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > gpurun_out/prof_r02${x}_source.csv.gz
done
python bench.py --objects 200000 --steps 2 --warmup 1 --no-cpu --gp1d-objects 0 > gpurun_out/plain_launches.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --objects 200000 --steps 2 --warmup 1 --no-cpu --gp1d-objects 0 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log; tail -1 gpurun_out/plain_launches.log | cut -c1-300
