#!/bin/bash
# ncu evidence of the current build: launch list, single-pass DRAM traffic, one --set full capture per kernel
set -x
mkdir -p gpurun_out
TAG=${1:-r02_a}
ncu --metrics gpu__time_duration.sum --clock-control none -s 36 -c 24 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/profile_step.py 12 > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_" -s 36 -c 12 --csv --log-file gpurun_out/traffic_$TAG.csv python scripts/profile_step.py 10 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_" -s 36 -c 6 -o gpurun_out/prof_$TAG -f python scripts/profile_step.py 8 > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
ls -la gpurun_out/prof_$TAG.ncu-rep
