import time, numpy as np, torch, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvom_b200 import Gvom, synth
P = synth.params_tuple()
g = Gvom(*P)
frames = [synth.frame(i, 128, 2048) for i in range(8)]
pinned = [torch.from_numpy(f[0]).pin_memory() for f in frames]
for mode in ("pageable", "pinned", "device"):
    ts = []
    for it in range(40):
        pc, ego, T = frames[it % 8]
        src = pc if mode == "pageable" else pinned[it % 8] if mode == "pinned" else None
        if src is None:
            src = pinned[it % 8].cuda(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        g.Process_pointcloud(src, ego, T)
        out = g.combine_maps()
        t1 = time.perf_counter()
        ts.append(t1 - t0)
    print(mode, "p50 ms", np.median(ts[10:]) * 1e3, "min", min(ts) * 1e3, "stage_copy_ms", g.stage_times()["stage_copy_host"])
g.set_profiling(True)
for it in range(3):
    pc, ego, T = frames[it]
    g.Process_pointcloud(pinned[it], ego, T); g.combine_maps()
    print(g.stage_times())
print(g.stats())
