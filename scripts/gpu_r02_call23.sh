#!/bin/bash
# after the removal of the superseded multi-GPU exchanges: full GPU suite on a 2-GPU box (incl. the torchrun checks), smoke, N=1 and N=2 bench
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_r02_c23_n1.json 2> gpurun_out/bench_r02_c23_n1.err
for tag in mirror bulk generic; do
  unset GVOM_VARIANT GVOM_MULTI_MIRROR
  [ $tag = bulk ] && export GVOM_VARIANT=32
  [ $tag = generic ] && export GVOM_MULTI_MIRROR=0
  timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29831 bench.py --gpus 2 --steps 100 --warmup 20 > gpurun_out/bench_r02_c23_n2_$tag.json 2> gpurun_out/bench_r02_c23_n2_$tag.err
done
python - <<PY
import json
for tag in ("n1","n2_mirror","n2_bulk","n2_generic"):
    try:
        d=json.loads(open("gpurun_out/bench_r02_c23_%s.json" % tag).read().strip().splitlines()[-1])
        print(tag, {k:d[k] for k in ("value","ms_per_step","p50_latency_ms","gpu_launches_per_step")}, d["io"]["exchange"], d.get("parity_check",{}).get("ok"), d["e2e"]["value"], d["e2e"]["p50_latency_ms"])
        print({k:round(1e3*v,1) for k,v in d["stage_ms"].items() if v})
    except Exception as e: print(tag, "ERR", e)
PY
