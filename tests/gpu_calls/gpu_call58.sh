#!/bin/bash
# K3 with shared (not inlined) team helpers, tight staging sizes and fused line-fit reductions against the previous build:
# time per pass and bit-comparison of every row at 200 k objects, then the parity tests of the passes that changed.
set -x
timeout 70 python tests/gpu_k3_variants.py 200000 old=variants/lib_k3old.so new=mallorn_astrophysics_b200/libmallorn_b200.so:1 new_loose=variants/lib_k3new_loose.so:0 inline_tight=variants/lib_k3inline.so:1 2>&1 | tail -6
timeout 60 python -m pytest tests/test_gpu_parity.py tests/test_gpu_research.py tests/test_gpu_frames.py -m gpu -x -q 2>&1 | tail -3
