#!/bin/bash
# The 10,178-object configs through the DataFrame API with the native factorisation of the frame's Arrow strings
# (store.frame_columns -> libmallorn_ingest.so), then the default C4 line in the final bench.py format.
set -x
for c in c2 c1 c3; do
  timeout 120 python bench.py --config $c > gpurun_out/bench_r02b_$c.json 2> gpurun_out/bench_r02b_$c.err; echo rc=$?; tail -1 gpurun_out/bench_r02b_$c.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_r02b_$c.json').read().strip().splitlines()[-1]); print('$c', d['value'], d['frames'])"
done
timeout 170 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r02b_1m.json 2> gpurun_out/bench_r02b_1m.err; echo rc=$?; tail -1 gpurun_out/bench_r02b_1m.err; cut -c1-400 gpurun_out/bench_r02b_1m.json
