#!/usr/bin/env python3
"""ncu counter CSV (--metrics ... --csv --log-file) -> per-kernel, per-launch averages merged into
profiles/kernel_counters_rNN.json under a configuration key (bench.py reads it for the issue / atomic bounds and
roofline.traffic):   python scripts/ncu_counters.py gpurun_out/traffic_r02_b.csv os1_128 profiles/kernel_counters_r02.json"""
import csv
import json
import os
import sys

src, config, dst = sys.argv[1], sys.argv[2], sys.argv[3]
NAMES = {"dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "gpu__time_duration.sum": "time_ns",
         "smsp__inst_executed.sum": "inst_executed", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum": "red_sectors",
         "l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum": "atom_sectors",
         "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
         "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
         "lts__t_bytes.sum": "l2_bytes"}
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hi]
ki, mi, vi = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
acc = {}
for r in rows[hi + 1:]:
    if len(r) != len(H) or r[mi] not in NAMES:
        continue
    k = r[ki].split("(")[0].replace("void ", "").strip()
    acc.setdefault(k, {}).setdefault(NAMES[r[mi]], []).append(float(r[vi].replace(",", "")))
out = json.load(open(dst)) if os.path.exists(dst) else {}
out[config] = {k: dict({m: sum(v) / len(v) for m, v in d.items()}, launches=len(next(iter(d.values())))) for k, d in acc.items()}
out.setdefault("_source", {})[config] = os.path.basename(src)
json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
print(json.dumps(out[config], indent=1))
