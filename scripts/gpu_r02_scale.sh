#!/bin/bash
# bench at N GPUs (row-sharded default) + optional legacy A/B
N=${1:-4}
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/multi_rank_check.py p2p grid256 2>&1 | grep -E "MULTI_RANK_OK|Error|error|assert|Traceback" | head -8 | tee gpurun_out/multi_rank_check_r02_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus $N --steps 100 --warmup 20 > gpurun_out/bench_r02_n${N}_rows.json 2> gpurun_out/bench_r02_n${N}_rows.err
tail -3 gpurun_out/bench_r02_n${N}_rows.err
if [ "$2" = "ab" ]; then
GVOM_MULTI_ROWS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus $N --steps 100 --warmup 20 > gpurun_out/bench_r02_n${N}_legacy.json 2> gpurun_out/bench_r02_n${N}_legacy.err
fi
python - <<PY
import json
for tag in ("rows","legacy"):
    try:
        d=json.loads(open("gpurun_out/bench_r02_n${N}_%s.json" % tag).read().strip().splitlines()[-1])
        print(tag, {k:d[k] for k in ("value","ms_per_step","p50_latency_ms","gpu_launches_per_step")}, d["io"]["exchange"], d.get("parity_check",{}).get("ok"))
        print({k:round(1e3*v,1) for k,v in d["stage_ms"].items() if v}); print(d["e2e"])
    except Exception as e: print(tag, "ERR", e)
PY
