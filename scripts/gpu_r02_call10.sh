#!/bin/bash
set -x
mkdir -p gpurun_out
for n in 1 2 4; do timeout 300 python scripts/mirror_probe.py $n 12 2>&1 | tail -1; done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_bytes.sum --clock-control none -k regex:"k_push_scan|k_mirror_args|k_merge_rows_ind|k_merge_cells2_ind|k_rows" -s 40 -c 30 --csv --log-file gpurun_out/mirror_probe_ncu.csv python scripts/mirror_probe.py 2 8 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/mirror_probe_ncu.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
agg={}
for r in rows[1:]:
    agg.setdefault((r[ii],r[ki][:28]),{})[r[mi]]=r[vi]
for k,v in list(agg.items())[:30]: print(k, v)
PY
