#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02_c6_tests.log; tail -6 gpurun_out/r02_c6_tests.log
timeout 300 python scripts/graph_probe.py 300 > gpurun_out/graph_probe_r02.jsonl 2> gpurun_out/graph_probe_r02.err; cat gpurun_out/graph_probe_r02.jsonl; tail -3 gpurun_out/graph_probe_r02.err
bash scripts/gpu_r02_ncu.sh r02_b os1_128 full > /dev/null 2>&1
python scripts/ncu_counters.py gpurun_out/traffic_r02_b.csv os1_128 profiles/kernel_counters_r02.json > /dev/null
cp profiles/kernel_counters_r02.json gpurun_out/kernel_counters_r02.json
for ch in 4 8; do
GVOM_CHUNKS=$ch timeout 300 python bench.py --steps 100 --warmup 20 > gpurun_out/bench_r02_c6_ch$ch.json 2> gpurun_out/bench_r02_c6_ch$ch.err
tail -2 gpurun_out/bench_r02_c6_ch$ch.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02_c6_ch$ch.json").read().strip().splitlines()[-1])
print("chunks $ch", {k:d[k] for k in ("value","ms_per_step","p50_latency_ms")}, d["e2e"]["p50_latency_ms"], d.get("e2e_variants_p50_ms"))
print({k:(v["bound"], round(v["frac"],3)) for k,v in d["rooflines"].items() if "frac" in v}, d["step_roofline"]["frac"], d.get("cpu_baseline_cudasim",{}).get("value"))
PY
done
