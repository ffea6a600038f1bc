#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_parity.py -x -q 2>&1 | tail -15 > gpurun_out/r02_c3_parity.log
cat gpurun_out/r02_c3_parity.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_c3_tests.log
tail -15 gpurun_out/r02_c3_tests.log
for v in 0 4; do
GVOM_VARIANT=$v timeout 300 python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench_r02_c3_v$v.json 2> gpurun_out/bench_r02_c3_v$v.err
tail -3 gpurun_out/bench_r02_c3_v$v.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02_c3_v$v.json").read().strip().splitlines()[-1])
print("variant $v", {k:d[k] for k in ("value","ms_per_step","p50_latency_ms","gpu_launches_per_step")}, d["e2e"]); print({k:round(1e3*v,1) for k,v in d["stage_ms"].items() if v})
PY
done
GVOM_VARIANT=0 timeout 300 python bench.py --config long_range --steps 60 --warmup 20 > gpurun_out/bench_r02_c3_long_range.json 2> gpurun_out/bench_r02_c3_long_range.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02_c3_long_range.json").read().strip().splitlines()[-1])
print("long_range", {k:d[k] for k in ("value","ms_per_step")}); print({k:round(1e3*v,1) for k,v in d["stage_ms"].items() if v})
PY
