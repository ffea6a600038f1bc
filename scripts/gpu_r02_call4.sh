#!/bin/bash
# sanitizer passes over the new kernels + e2e input-path A/B
set -x
mkdir -p gpurun_out
python scripts/sanitizer_case.py > gpurun_out/r02_sanitizer_plain.log 2>&1; tail -2 gpurun_out/r02_sanitizer_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitizer_case.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZER_CASES_OK|case xy" gpurun_out/r02_sanitizer_$tool.log | tail -6
done
for m in zero dma; do
GVOM_H2D=$m timeout 300 python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench_r02_c4_$m.json 2> gpurun_out/bench_r02_c4_$m.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02_c4_$m.json").read().strip().splitlines()[-1])
print("$m", {k:d[k] for k in ("value","ms_per_step","p50_latency_ms")}, d["e2e"], d.get("e2e_variants_p50_ms"))
PY
done
