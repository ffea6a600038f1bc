#!/usr/bin/env python3
"""Stall samples / executed instructions in buckets of SASS lines: python scripts/ncu_lines.py rep kernel_regex [bucket]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 100
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(H)]
seen, uniq = set(), []
for r in data:
    if r[0] in seen: break
    seen.add(r[0]); uniq.append(r)
data = uniq
si, ni, ii = H.index("Source"), H.index("# Samples"), H.index("Instructions Executed")
I = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0
tot = sum(I(r[ni]) for r in data); tex = sum(I(r[ii]) for r in data)
for b in range(0, len(data), B):
    ch = data[b:b + B]
    s = sum(I(r[ni]) for r in ch); e = sum(I(r[ii]) for r in ch)
    ops = {}
    for r in ch:
        op = r[si].split()[0] if not r[si].startswith("@") else r[si].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:5]
    print(f"{b:5d} samples {s:5d} {100*s/max(tot,1):5.1f}%  exec {e:9d} {100*e/max(tex,1):5.1f}%  {top}")
