#!/bin/bash
# sanitizer passes over the multi-GPU kernels and the bulk-copy staging of host clouds
set -x
mkdir -p gpurun_out
python scripts/sanitizer_mirror.py > gpurun_out/r02_sanitizer_mirror_plain.log 2>&1; tail -2 gpurun_out/r02_sanitizer_mirror_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitizer_mirror.py > gpurun_out/r02_sanitizer_mirror_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZER_MIRROR_OK|mirror case" gpurun_out/r02_sanitizer_mirror_$tool.log | tail -4
done
for tool in memcheck racecheck; do
  timeout 500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitizer_case.py > gpurun_out/r02_sanitizer2_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZER_CASES_OK|case xy" gpurun_out/r02_sanitizer2_$tool.log | tail -5
done
