#!/bin/bash
set -x
mkdir -p gpurun_out
for n in 8; do timeout 600 python scripts/mirror_probe.py $n 10 2>&1 | tail -1; done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_push_scan|k_mirror_args|k_merge_rows|k_merge_cells2|k_rows_known|k_surface" -s 160 -c 64 --csv --log-file gpurun_out/mirror_probe_ncu8.csv python scripts/mirror_probe.py 8 6 > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/mirror_probe_ncu8.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
agg={}
for r in rows[1:]:
    agg.setdefault((r[ii],r[ki].split("(")[0]),{})[r[mi]]=float(r[vi].replace(",",""))
by=collections.defaultdict(list)
for (i,k),v in agg.items(): by[k].append(v)
for k,vs in by.items():
    n=len(vs); m=lambda key: sum(v.get(key,0) for v in vs)/n
    print(k, n, "us", round(m("gpu__time_duration.sum")/1e3,1), "dramR MB", round(m("dram__bytes_read.sum")/1e6,1), "L2 MB", round(m("lts__t_bytes.sum")/1e6,1), "inst", int(m("smsp__inst_executed.sum")), "warps%", round(m("sm__warps_active.avg.pct_of_peak_sustained_active"),1))
PY
