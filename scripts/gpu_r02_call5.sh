#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -25 > gpurun_out/r02_c5_multi.log
cat gpurun_out/r02_c5_multi.log
