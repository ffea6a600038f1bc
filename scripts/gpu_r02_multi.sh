#!/bin/bash
# multi-GPU: parity of every exchange on real peers (logs kept), then the bench at this world size
N=${1:-2}
set -x
mkdir -p gpurun_out
port=29600
: > gpurun_out/multi_rank_check_r02_n$N.log
for spec in "nccl" "p2p grid256" "p2p grid256 late" "p2p" "p2p late" "p2p sharded"; do
  port=$((port+1))
  echo "=== torchrun x$N tests/multi_rank_check.py $spec" >> gpurun_out/multi_rank_check_r02_n$N.log
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port tests/multi_rank_check.py $spec 2>&1 | grep -E "MULTI_RANK_OK|Error|error|assert|Traceback" | head -8 >> gpurun_out/multi_rank_check_r02_n$N.log
done
cat gpurun_out/multi_rank_check_r02_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus $N --steps 100 --warmup 20 > gpurun_out/bench_r02_n${N}_rows.json 2> gpurun_out/bench_r02_n${N}_rows.err
tail -3 gpurun_out/bench_r02_n${N}_rows.err
GVOM_MULTI_ROWS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus $N --steps 100 --warmup 20 > gpurun_out/bench_r02_n${N}_legacy.json 2> gpurun_out/bench_r02_n${N}_legacy.err
tail -3 gpurun_out/bench_r02_n${N}_legacy.err
python - <<PY
import json
for tag in ("rows","legacy"):
    try:
        d=json.loads(open("gpurun_out/bench_r02_n${N}_%s.json" % tag).read().strip().splitlines()[-1])
        print(tag, {k:d[k] for k in ("value","ms_per_step","p50_latency_ms")}, d["io"]["exchange"], d.get("parity_check"))
        print({k:round(1e3*v,1) for k,v in d["stage_ms"].items() if v})
    except Exception as e: print(tag, "ERR", e)
PY
