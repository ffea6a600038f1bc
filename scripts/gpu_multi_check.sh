#!/bin/bash
# multi-GPU check at N ranks (gpurun --gpus N): in-process parity, torchrun parity incl. a late-joining rank (log kept), bench.py --gpus N
N=${1:-2}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -k "mirrored" 2>&1 | tail -3
port=29740
LOG=gpurun_out/multi_rank_check_r02_mirror_n$N.log
: > $LOG
for spec in "p2p grid256" "p2p grid256 late"; do
  port=$((port+1))
  echo "=== torchrun x$N tests/multi_rank_check.py $spec" >> $LOG
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port tests/multi_rank_check.py $spec 2>&1 | grep -E "MULTI_RANK_OK|Error|error|assert|Traceback" | head -8 >> $LOG
done
cat $LOG
for tag in mirror; do
  port=$((port+1))
  timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 100 --warmup 20 > gpurun_out/bench_r02_c12_n${N}_$tag.json 2> gpurun_out/bench_r02_c12_n${N}_$tag.err
  tail -3 gpurun_out/bench_r02_c12_n${N}_$tag.err
done
python - <<PY
import json
for tag in ("n${N}_mirror",):
    try:
        d=json.loads(open("gpurun_out/bench_r02_c12_%s.json" % tag).read().strip().splitlines()[-1])
        print(tag, {k:d[k] for k in ("value","ms_per_step","p50_latency_ms","gpu_launches_per_step")}, d["io"]["exchange"], d.get("parity_check",{}).get("ok"), d["e2e"]["value"], d["e2e"]["p50_latency_ms"])
        print({k:round(1e3*v,1) for k,v in d["stage_ms"].items() if v})
    except Exception as e: print(tag, "ERR", e)
PY
