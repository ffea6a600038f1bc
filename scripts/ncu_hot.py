#!/usr/bin/env python3
"""Top SASS lines by stall samples for one kernel: python scripts/ncu_hot.py rep kernel_regex [N]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(H)]
# ncu lists every instantiation; keep the first block of addresses only
seen, uniq = set(), []
for r in data:
    if r[0] in seen: break
    seen.add(r[0]); uniq.append(r)
data = uniq
si, ni, ii = H.index("Source"), H.index("# Samples"), H.index("Instructions Executed")
stall = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
I = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0
tot = sum(I(r[ni]) for r in data); texec = sum(I(r[ii]) for r in data)
print(f"samples {tot}  warp-instructions {texec}  sass lines {len(data)}")
agg = {}
for r in data:
    for i in stall: agg[H[i][6:]] = agg.get(H[i][6:], 0) + I(r[i])
print("stall mix:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for k, r in sorted(enumerate(data), key=lambda kr: -I(kr[1][ni]))[:N]:
    st = sorted(((I(r[i]), H[i][6:]) for i in stall), reverse=True)[:2]
    print(f"{k:4d} {I(r[ni]):6d} {100*I(r[ni])/max(tot,1):5.1f}% ex={I(r[ii]):8d} {r[si][:72]:72s} {st}")
