#!/bin/bash
set -x
timeout 600 python tests/gpu_overlap2.py 200000
timeout 600 python tests/gpu_overlap2.py 1000000
