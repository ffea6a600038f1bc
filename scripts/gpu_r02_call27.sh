#!/bin/bash
set -x
mkdir -p gpurun_out
for i in 1 2; do
  timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python scripts/sanitizer_mirror.py > gpurun_out/r02_sanitizer_mirror_racecheck$i.log 2>&1
  echo "racecheck rc=$?"; grep -E "MISMATCH|RACECHECK SUMMARY|SANITIZER_MIRROR_OK|mirror case" gpurun_out/r02_sanitizer_mirror_racecheck$i.log | tail -4
done
for i in 1 2 3 4 5 6; do python scripts/sanitizer_mirror.py 2>&1 | grep -E "MISMATCH|SANITIZER_MIRROR_OK" | tail -2; done
