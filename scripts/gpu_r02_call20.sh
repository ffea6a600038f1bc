#!/bin/bash
set -x
mkdir -p gpurun_out
GVOM_VARIANT=32 timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -k "mirrored" 2>&1 | tail -3
for v in 0 32; do
GVOM_VARIANT=$v ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_bytes.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_push_scan" -s 8 -c 8 --csv --log-file gpurun_out/push_ncu_v$v.csv python scripts/mirror_probe.py 2 8 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/push_ncu_v$v.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value")
d={}
for r in rows[1:]: d.setdefault((r[ki].split("(")[0], r[mi]), []).append(float(r[vi].replace(",","")))
for k,v in d.items(): print("variant $v", k, round(sum(v)/len(v),1))
PY
done
