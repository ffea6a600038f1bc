#!/bin/bash
# ncu durations of the shipped multi-GPU kernels, 2 and 8 ranks played on one GPU (no NVLink, no rank skew)
set -x
mkdir -p gpurun_out
for n in 2 8; do
timeout 300 python scripts/mirror_probe.py $n 10 2>&1 | tail -1 > gpurun_out/mirror_probe_stage_n$n.txt; cat gpurun_out/mirror_probe_stage_n$n.txt
skip=$((n*20)); cnt=$((n*9))
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_push_scan|k_mirror_args|k_merge_rows|k_merge_cells2|k_rows_known|k_surface|k_rows_deliver|k_scan" -s $((n*44)) -c $((n*11)) --csv --log-file gpurun_out/mirror_probe_ncu_final_n$n.csv python scripts/mirror_probe.py $n 6 > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/mirror_probe_ncu_final_n$n.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
agg={}
for r in rows[1:]:
    agg.setdefault((r[ii],r[ki].split("(")[0]),{})[r[mi]]=float(r[vi].replace(",",""))
by=collections.defaultdict(list)
for (i,k),v in agg.items(): by[k].append(v)
for k,vs in by.items():
    m=lambda key: sum(v.get(key,0) for v in vs)/len(vs)
    print("n=$n", k, len(vs), "us", round(m("gpu__time_duration.sum")/1e3,1), "dramR MB", round(m("dram__bytes_read.sum")/1e6,1), "L2 MB", round(m("lts__t_bytes.sum")/1e6,1), "inst", int(m("smsp__inst_executed.sum")), "warps%", round(m("sm__warps_active.avg.pct_of_peak_sustained_active"),1))
PY
done
