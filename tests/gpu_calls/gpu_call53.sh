#!/bin/bash
set -x
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-300
python bench.py --objects 200000 --steps 2 --warmup 3 --no-cpu --gp1d-objects 0 2> gpurun_out/b200k.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k:round(v['ms_per_step'],1) for k,v in d['passes'].items()}, d['roofline']['frac'], d['roofline_dominant']['frac'])"
