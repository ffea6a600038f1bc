#!/bin/bash
set -x
mkdir -p gpurun_out
for tool in racecheck memcheck synccheck; do
timeout 400 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitizer_mirror.py > gpurun_out/r02_sanitizer_mirror_$tool.log 2>&1
echo "$tool rc=$?"; grep -E "MISMATCH|RACECHECK SUMMARY|ERROR SUMMARY|SANITIZER_MIRROR_OK|mirror case|Assert" gpurun_out/r02_sanitizer_mirror_$tool.log | tail -4
done
