#!/bin/bash
# final evidence of the round on one GPU: ncu launch lists + counters of the shipped build (all three configurations), one
# --set full capture, then pytest -m gpu, smoke, the three bench configurations and the reference arm
set -x
mkdir -p gpurun_out
for c in os1_128 long_range dense; do
  full=""; [ $c = os1_128 ] && full=full
  bash scripts/gpu_ncu.sh r02_final_$c $c $full > /dev/null 2>&1
  python scripts/ncu_counters.py gpurun_out/traffic_r02_final_$c.csv $c profiles/kernel_counters_r02.json > /dev/null
done
cp profiles/kernel_counters_r02.json gpurun_out/kernel_counters_r02.json
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_final_tests.log; tail -3 gpurun_out/r02_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_r02_final_n1.json 2> gpurun_out/bench_r02_final_n1.err; tail -2 gpurun_out/bench_r02_final_n1.err
timeout 300 python bench.py --config long_range --steps 100 --warmup 20 > gpurun_out/bench_r02_final_long_range.json 2> gpurun_out/bench_r02_final_long_range.err
timeout 300 python bench.py --config dense --steps 16 --warmup 4 > gpurun_out/bench_r02_final_dense.json 2> gpurun_out/bench_r02_final_dense.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r02_final_reference.json 2> gpurun_out/bench_r02_final_reference.err
python - <<PY
import json
for t in ("n1","long_range","dense"):
    try:
        d=json.loads(open("gpurun_out/bench_r02_final_%s.json" % t).read().strip().splitlines()[-1])
        print(t, {k:d[k] for k in ("value","ms_per_step","p50_latency_ms")}, d["e2e"]["value"], d["e2e"]["p50_latency_ms"])
        print("  ", {k:round(1e3*v,1) for k,v in d["stage_ms"].items() if v})
        print("  ", {k:(v["bound"], round(v["frac"],3)) for k,v in d["rooflines"].items() if "frac" in v}, round(d["step_roofline"]["frac"],3))
        if t == "n1": print("  ", d.get("e2e_variants_p50_ms"), d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline_cudasim"))
    except Exception as e: print(t, "ERR", e)
d=json.loads(open("gpurun_out/bench_r02_final_reference.json").read().strip().splitlines()[-1]); print("reference", {k:d.get(k) for k in ("value","ms_per_step","value_p50","impl")})
PY
