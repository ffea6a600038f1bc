#!/bin/bash
# full validation of the current build on one GPU: pytest -m gpu, smoke, the three bench configurations, reference arm
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02_c14_tests.log; tail -4 gpurun_out/r02_c14_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 400 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_r02_c14_n1.json 2> gpurun_out/bench_r02_c14_n1.err; tail -2 gpurun_out/bench_r02_c14_n1.err
timeout 300 python bench.py --config long_range --steps 100 --warmup 20 > gpurun_out/bench_r02_c14_long_range.json 2> gpurun_out/bench_r02_c14_long_range.err
timeout 300 python bench.py --config dense --steps 16 --warmup 4 > gpurun_out/bench_r02_c14_dense.json 2> gpurun_out/bench_r02_c14_dense.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r02_c14_reference.json 2> gpurun_out/bench_r02_c14_reference.err; tail -2 gpurun_out/bench_r02_c14_reference.err
python - <<PY
import json
for t in ("n1","long_range","dense"):
    try:
        d=json.loads(open("gpurun_out/bench_r02_c14_%s.json" % t).read().strip().splitlines()[-1])
        print(t, {k:d[k] for k in ("value","ms_per_step","p50_latency_ms")}, d["e2e"]["value"], d["e2e"]["p50_latency_ms"])
        print("  ", {k:round(1e3*v,1) for k,v in d["stage_ms"].items() if v})
        print("  ", {k:(v["bound"], round(v["frac"],3)) for k,v in d["rooflines"].items() if "frac" in v}, round(d["step_roofline"]["frac"],3))
        if t == "n1": print("  ", d.get("e2e_variants_p50_ms"), d.get("cpu_baseline"))
    except Exception as e: print(t, "ERR", e)
d=json.loads(open("gpurun_out/bench_r02_c14_reference.json").read().strip().splitlines()[-1]); print("reference", {k:d.get(k) for k in ("value","ms_per_step","value_p50","impl")})
PY
