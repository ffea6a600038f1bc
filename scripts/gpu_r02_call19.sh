#!/bin/bash
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_merge_cells2_rows|k_surface_maps2|k_merge_rows_ind" -s 60 -c 3 -o gpurun_out/prof_r02_mirror8 -f python scripts/mirror_probe.py 8 6 > gpurun_out/ncu_r02_mirror8.log 2>&1
tail -3 gpurun_out/ncu_r02_mirror8.log
ls -la gpurun_out/prof_r02_mirror8.ncu-rep
