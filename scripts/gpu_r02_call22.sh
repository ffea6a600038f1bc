#!/bin/bash
# third build of the cell merge (8 lanes per cell): full GPU suite + N=1 bench A/B
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for v in 0 1; do
GVOM_VARIANT=$v timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_r02_c22_v$v.json 2> gpurun_out/bench_r02_c22_v$v.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02_c22_v$v.json").read().strip().splitlines()[-1])
print("variant $v", {k:d[k] for k in ("value","ms_per_step","p50_latency_ms")}, d["e2e"]["value"], d["e2e"]["p50_latency_ms"])
print({k:round(1e3*v,1) for k,v in d["stage_ms"].items() if v})
PY
done
