#!/usr/bin/env python3
"""Per-stage CUDA-event times (library profiling events) for three workload sizes: how much of a stage is a
fixed floor (launch + event + dependent round trips) and how much scales with the work."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gvom_b200 import Gvom, synth

for name, P, beams, cols, wall in (("tiny 64x64x16, 4k pts", synth.params_tuple(xy_size=64, z_size=16, robot_radius=2.0), 16, 256, 9.0),
                                   ("OS1-64 65k pts, 256x256x64", synth.params_tuple(), 64, 1024, 45.0),
                                   ("OS1-128 262k pts, 256x256x64", synth.params_tuple(), 128, 2048, 45.0)):
    g = Gvom(*P)
    fr = [synth.frame(i, beams, cols, wall_radius=wall) for i in range(4)]
    dev = [torch.from_numpy(f[0]).cuda() for f in fr]
    g.set_profiling(True)
    acc, n = {}, 0
    for i in range(40):
        k = i % 4
        g.Process_pointcloud(dev[k], fr[k][1], fr[k][2]); g.combine_maps(device_outputs=True)
        if i >= 10:
            n += 1
            for kk, vv in g.stage_times().items():
                acc[kk] = acc.get(kk, 0.0) + vv
    print(json.dumps({"workload": name, "stage_us": {k: round(1e3 * v / n, 2) for k, v in acc.items() if v}}), flush=True)
