#!/bin/bash
# S2 pushes the scan itself (occupied marker from S1): full GPU suite, N=1 bench, then N ranks: push inside S2 vs separate push kernel
N=${1:-2}
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_r02_c15_n1.json 2> gpurun_out/bench_r02_c15_n1.err
tail -2 gpurun_out/bench_r02_c15_n1.err
port=29760
for spec in "p2p grid256" "p2p grid256 late"; do
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port tests/multi_rank_check.py $spec 2>&1 | grep -E "MULTI_RANK_OK|Error|error|assert|Traceback" | head -8
done
for tag in s2push pushkernel; do
  [ $tag = pushkernel ] && export GVOM_VARIANT=8
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 100 --warmup 20 > gpurun_out/bench_r02_c15_n${N}_$tag.json 2> gpurun_out/bench_r02_c15_n${N}_$tag.err
  tail -3 gpurun_out/bench_r02_c15_n${N}_$tag.err
done
python - <<PY
import json
for tag in ("n1", "n${N}_s2push","n${N}_pushkernel"):
    try:
        d=json.loads(open("gpurun_out/bench_r02_c15_%s.json" % tag).read().strip().splitlines()[-1])
        print(tag, {k:d[k] for k in ("value","ms_per_step","p50_latency_ms","gpu_launches_per_step")}, d["io"]["exchange"], d.get("parity_check",{}).get("ok"), d["e2e"]["value"], d["e2e"]["p50_latency_ms"])
        print({k:round(1e3*v,1) for k,v in d["stage_ms"].items() if v})
    except Exception as e: print(tag, "ERR", e)
PY
