#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into a markdown table: python scripts/ncu_summary.py rep [--stalls kernel]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H = rows[0]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram rd MB"),
        ("dram__bytes_write.sum", "dram wr MB"), ("lts__t_bytes.sum", "L2 MB"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"), ("launch__registers_per_thread", "regs"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "RED sectors"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %")]
idx = [(H.index(k), n) for k, n in want if k in H]
units = rows[1]
print("| " + " | ".join(n for _, n in idx) + " |")
print("|" + "---|" * len(idx))
for r in rows[2:]:
    cells = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = "`" + v.split("(")[0].replace("void ", "") + "`"
        else:
            try:
                f = float(v.replace(",", ""))
                if "MB" in n and units[i].lower().startswith("byte"):
                    f /= 1e6
                if n == "us" and units[i] in ("ns", "nsecond"):
                    f /= 1e3
                v = f"{f:.2f}" if f < 1000 else f"{f:.0f}"
            except ValueError:
                pass
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
if len(sys.argv) > 3 and sys.argv[2] == "--stalls":
    det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(det)))
    H = rows[0]
    ki, mi, vi, si = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Section Name")
    for r in rows[1:]:
        if sys.argv[3] in r[ki] and r[si] in ("Warp State Statistics", "Scheduler Statistics", "Memory Workload Analysis", "Occupancy"):
            print(f"{r[si][:12]:12s} {r[mi]:55s} {r[vi]}")
