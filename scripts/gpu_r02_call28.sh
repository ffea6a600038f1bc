#!/bin/bash
set -x
mkdir -p gpurun_out
python scripts/sanitizer_mirror.py 2>&1 | grep -E "MISMATCH|SANITIZER_MIRROR_OK|Assert|slot" | tail -3
timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python scripts/sanitizer_mirror.py > gpurun_out/r02_sanitizer_mirror_racecheck3.log 2>&1
echo "racecheck rc=$?"; grep -E "MISMATCH|RACECHECK SUMMARY|SANITIZER_MIRROR_OK|mirror case|Assert|slot|  voxel" gpurun_out/r02_sanitizer_mirror_racecheck3.log | tail -14
