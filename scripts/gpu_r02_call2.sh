#!/bin/bash
# round 2, GPU call 2: first run of the two-kernel scan path (S1 k_scan_points, S2 k_scan_cells)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_parity.py -x -q -k "tiny or small" 2>&1 | tail -15 > gpurun_out/r02_c2_small.log
cat gpurun_out/r02_c2_small.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_c2_tests.log
tail -15 gpurun_out/r02_c2_tests.log
timeout 300 python bench.py --steps 100 --warmup 20 > gpurun_out/bench_r02_c2_n1.json 2> gpurun_out/bench_r02_c2_n1.err
tail -3 gpurun_out/bench_r02_c2_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02_c2_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","p50_latency_ms","gpu_launches_per_step")}, d["e2e"]); print(d["stage_ms"])
PY
