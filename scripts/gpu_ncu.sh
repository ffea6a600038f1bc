#!/bin/bash
# ncu evidence of the current build: launch list, single-pass counters, one --set full capture per kernel
#   bash scripts/gpu_ncu.sh TAG [config] [full]
set -x
mkdir -p gpurun_out
TAG=${1:-r02_a}
CFG=${2:-os1_128}
SKIP=36; [ "$CFG" = "dense" ] && SKIP=18
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 24 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/profile_step.py 12 --config $CFG > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none -k regex:"k_" -s $SKIP -c 12 --csv --log-file gpurun_out/traffic_$TAG.csv python scripts/profile_step.py 10 --config $CFG > /dev/null 2>&1
if [ "$3" = "full" ]; then
ncu --set full --clock-control none --import-source on -k regex:"k_" -s $SKIP -c 6 -o gpurun_out/prof_$TAG -f python scripts/profile_step.py 8 --config $CFG > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
fi
