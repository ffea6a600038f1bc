#!/bin/bash
N=${1:-2}
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -k "row_sharded or partial_finish" 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/multi_rank_check.py p2p grid256 2>&1 | grep -E "MULTI_RANK_OK|Error|error|assert|Traceback" | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus $N --steps 100 --warmup 20 > gpurun_out/bench_r02_n${N}_rows2.json 2> gpurun_out/bench_r02_n${N}_rows2.err
tail -3 gpurun_out/bench_r02_n${N}_rows2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02_n${N}_rows2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","p50_latency_ms","gpu_launches_per_step")}, d["io"]["exchange"], d.get("parity_check",{}).get("ok"))
print({k:round(1e3*v,1) for k,v in d["stage_ms"].items() if v}); print(d["e2e"])
PY
