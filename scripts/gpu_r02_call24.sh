#!/bin/bash
set -x
mkdir -p gpurun_out
for v in 0 256; do
GVOM_VARIANT=$v timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_r02_c24_v$v.json 2> gpurun_out/bench_r02_c24_v$v.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02_c24_v$v.json").read().strip().splitlines()[-1])
print("variant $v", {k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"]["p50_latency_ms"], d["e2e"]["p50_device_ms"])
PY
done
GVOM_VARIANT=256 timeout 300 python -m pytest tests/test_cuda_parity.py -q -x -k "input_variants or golden" 2>&1 | tail -2
