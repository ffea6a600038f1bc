#!/bin/bash
# pipelined bench loop + tuned push / args kernels: N=1, then N=2 mirror vs partial-merge rows
N=${1:-2}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -k "mirrored" 2>&1 | tail -3
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_r02_c11_n1.json 2> gpurun_out/bench_r02_c11_n1.err
tail -2 gpurun_out/bench_r02_c11_n1.err
port=29720
for tag in mirror rows; do
  [ $tag = rows ] && export GVOM_MULTI_MIRROR=0
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 100 --warmup 20 > gpurun_out/bench_r02_c11_n${N}_$tag.json 2> gpurun_out/bench_r02_c11_n${N}_$tag.err
  tail -3 gpurun_out/bench_r02_c11_n${N}_$tag.err
done
python - <<PY
import json
for tag in ("n1", "n${N}_mirror","n${N}_rows"):
    try:
        d=json.loads(open("gpurun_out/bench_r02_c11_%s.json" % tag).read().strip().splitlines()[-1])
        print(tag, {k:d[k] for k in ("value","ms_per_step","p50_latency_ms","gpu_launches_per_step")}, d["io"]["exchange"], d.get("parity_check",{}).get("ok"), d["e2e"]["value"], d["e2e"]["p50_latency_ms"])
        print({k:round(1e3*v,1) for k,v in d["stage_ms"].items() if v})
    except Exception as e: print(tag, "ERR", e)
PY
