#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -k "mirrored" 2>&1 | tail -3
for n in 2 8; do
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_merge_cells2_rows|k_merge_rows_ind" -s $((n*8)) -c $((n*4)) --csv --log-file gpurun_out/cells_rows_ncu_n$n.csv python scripts/mirror_probe.py $n 6 > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/cells_rows_ncu_n$n.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
agg={}
for r in rows[1:]:
    agg.setdefault((r[ii],r[ki].split("(")[0]),{})[r[mi]]=float(r[vi].replace(",",""))
by=collections.defaultdict(list)
for (i,k),v in agg.items(): by[k].append(v)
for k,vs in by.items():
    m=lambda key: sum(v.get(key,0) for v in vs)/len(vs)
    print("n=$n", k, len(vs), "us", round(m("gpu__time_duration.sum")/1e3,1), "inst", int(m("smsp__inst_executed.sum")), "warps%", round(m("sm__warps_active.avg.pct_of_peak_sustained_active"),1))
PY
done
