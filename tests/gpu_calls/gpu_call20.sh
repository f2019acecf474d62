#!/bin/bash
set -x
MLRN_LIB=variants/lib_fitstats.so timeout 300 python tests/gpu_fit_stats.py bazin 1000000
MLRN_LIB=variants/lib_fitstats.so timeout 300 python tests/gpu_fit_stats.py powerlaw 1000000
MLRN_LIB=variants/lib_fitstats.so timeout 300 python tests/gpu_fit_stats.py bazin 200000
