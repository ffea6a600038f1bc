#!/bin/bash
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_push_scan" -s 8 -c 1 -o gpurun_out/prof_r02_push -f python scripts/mirror_probe.py 2 8 > gpurun_out/ncu_r02_push.log 2>&1
tail -2 gpurun_out/ncu_r02_push.log
