import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gvom_b200 import Gvom, synth
g = Gvom(*synth.params_tuple())
frames = [synth.frame(i, 128, 2048) for i in range(4)]
g.set_profiling(True)
for it in range(12):
    pc, ego, T = frames[it % 4]
    t0 = time.perf_counter(); g.Process_pointcloud(pc, ego, T); t1 = time.perf_counter(); g.combine_maps(); t2 = time.perf_counter()
    if it >= 8:
        st = g.stage_times()
        print(f"process {1e3*(t1-t0):.3f} combine {1e3*(t2-t1):.3f} ms | " + " ".join(f"{k}={v*1e3:.0f}us" for k, v in st.items()))
