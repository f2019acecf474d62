#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_properties.py tests/test_gpu_frames.py -m gpu -x -q 2>&1 | tail -4
python bench.py --objects 200000 --steps 2 --warmup 1 --no-cpu --gp1d-objects 0 2> gpurun_out/b200k.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k:round(v['ms_per_step'],1) for k,v in d['passes'].items()})"
timeout 600 python bench.py --config c3 2> gpurun_out/c3.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3', d['value'], d['ms_per_step'])"
