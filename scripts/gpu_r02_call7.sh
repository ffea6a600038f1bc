#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02_c7_tests.log; tail -4 gpurun_out/r02_c7_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_r02_c7_n1.json 2> gpurun_out/bench_r02_c7_n1.err; tail -2 gpurun_out/bench_r02_c7_n1.err
for c in long_range dense; do
bash scripts/gpu_r02_ncu.sh r02_$c $c > /dev/null 2>&1
python scripts/ncu_counters.py gpurun_out/traffic_r02_$c.csv $c profiles/kernel_counters_r02.json > /dev/null
done
cp profiles/kernel_counters_r02.json gpurun_out/kernel_counters_r02.json
timeout 300 python bench.py --config long_range --steps 100 --warmup 20 > gpurun_out/bench_r02_c7_long_range.json 2> gpurun_out/bench_r02_c7_long_range.err
timeout 300 python bench.py --config dense --steps 16 --warmup 4 > gpurun_out/bench_r02_c7_dense.json 2> gpurun_out/bench_r02_c7_dense.err
python - <<PY
import json
for t in ("n1","long_range","dense"):
    try:
        d=json.loads(open("gpurun_out/bench_r02_c7_%s.json" % t).read().strip().splitlines()[-1])
        print(t, {k:d[k] for k in ("value","ms_per_step","p50_latency_ms")}, d["e2e"]["value"], d["e2e"]["p50_latency_ms"])
        print("  ", {k:round(1e3*v,1) for k,v in d["stage_ms"].items() if v})
        print("  ", {k:(v["bound"], round(v["frac"],3)) for k,v in d["rooflines"].items() if "frac" in v}, round(d["step_roofline"]["frac"],3))
    except Exception as e: print(t, "ERR", e)
PY
