"""Stage times of the mirrored multi-GPU combine with N "ranks" played by one process on ONE GPU (all blocks in
local memory): separates kernel cost from NVLink / rank-skew effects.  usage: mirror_probe.py [nranks] [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import bench  # noqa: E402
from gvom_b200 import Gvom, synth  # noqa: E402
from test_multi_gpu import attach_mirrors, local_exchange_mirror  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
P = synth.params_tuple(buffer_size=2)
ranks = [Gvom(*P) for _ in range(n)]
blocks = attach_mirrors(ranks)
fr = [bench.frames(r) for r in range(n)]
dev = [[torch.from_numpy(f[0]).cuda() for f in fr[r]] for r in range(n)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
acc = {}
for g in ranks:
    g.set_profiling(True)
for i in range(steps):
    k = i % len(fr[0])
    flush.zero_()
    torch.cuda.synchronize()
    for r, g in enumerate(ranks):
        g.Process_pointcloud(dev[r][k], fr[r][k][1], fr[r][k][2])
    torch.cuda.synchronize()
    t_proc = [g.stage_times() for g in ranks]
    outs, _ = local_exchange_mirror(ranks, i + 1, blocks)
    t_comb = [g.stage_times() for g in ranks]
    if i >= 4:
        for key in ("scan_points", "scan_cells", "push_or_partial"):
            acc.setdefault(key, []).append(np.mean([t[key] for t in t_proc]))
        for key in ("merge_codes", "merge_cells", "rows_surface", "d2h"):
            acc.setdefault(key, []).append(np.mean([t[key] for t in t_comb]))
print(f"ranks {n}:", {k: round(1e3 * float(np.median(v)), 1) for k, v in acc.items()}, "us (push_or_partial = push kernel; merge_codes includes the flag exchange; rows_surface = known + surface; d2h = deliver)")
