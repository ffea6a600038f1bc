#!/bin/bash
# mirrored ring slots, first run: in-process parity (1 GPU) + N=1 bench of the refactored merge kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -k "mirrored or row_sharded" 2>&1 | tail -15
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_r02_c8_n1.json 2> gpurun_out/bench_r02_c8_n1.err
tail -2 gpurun_out/bench_r02_c8_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02_c8_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","p50_latency_ms")}, d["e2e"]["value"], d["e2e"]["p50_latency_ms"])
print({k:round(1e3*v,1) for k,v in d["stage_ms"].items() if v})
PY
