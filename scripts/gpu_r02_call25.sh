#!/bin/bash
set -x
mkdir -p gpurun_out
for sp in 0 30 45 60; do
GVOM_SPLIT_H2D=$sp timeout 300 python bench.py --steps 150 --warmup 20 --no-cpu-baseline > gpurun_out/bench_r02_c25_s$sp.json 2> gpurun_out/bench_r02_c25_s$sp.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02_c25_s$sp.json").read().strip().splitlines()[-1])
print("split $sp", {k:round(d[k],4) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), d["e2e"]["p50_latency_ms"], d["e2e"]["p50_device_ms"])
PY
done
GVOM_SPLIT_H2D=45 timeout 300 python -m pytest tests/test_cuda_parity.py -q -x -k "input_variants or golden" 2>&1 | tail -2
