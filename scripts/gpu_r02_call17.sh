#!/bin/bash
# source masks (row merge -> cell kernel) on/off at N ranks
N=${1:-8}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -k "mirrored" 2>&1 | tail -3
port=29800
for tag in masks nomasks; do
  [ $tag = nomasks ] && export GVOM_VARIANT=16
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 100 --warmup 20 > gpurun_out/bench_r02_c17_n${N}_$tag.json 2> gpurun_out/bench_r02_c17_n${N}_$tag.err
  tail -3 gpurun_out/bench_r02_c17_n${N}_$tag.err
done
python - <<PY
import json
for tag in ("n${N}_masks","n${N}_nomasks"):
    try:
        d=json.loads(open("gpurun_out/bench_r02_c17_%s.json" % tag).read().strip().splitlines()[-1])
        print(tag, {k:d[k] for k in ("value","ms_per_step","p50_latency_ms","gpu_launches_per_step")}, d.get("parity_check",{}).get("ok"), d["e2e"]["value"], d["e2e"]["p50_latency_ms"])
        print({k:round(1e3*v,1) for k,v in d["stage_ms"].items() if v})
    except Exception as e: print(tag, "ERR", e)
PY
