#!/usr/bin/env python3
"""CUDA-graph replay of the tick vs the PDL launch chain (gvom_graph_probe), at full size and for a
small scan:   python scripts/graph_probe.py [iters]  ->  one JSON line per case."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gvom_b200 import Gvom, synth  # noqa: E402
from gvom_b200._lib import GVOM_F64, check  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for name, beams, cols, P in (("os1_128 262,144 pts, 256x256x64", 128, 2048, synth.params_tuple()),
                             ("4,096 pts, 64x64x16", 16, 256, synth.params_tuple(xy_size=64, z_size=16, robot_radius=2.0))):
    g = Gvom(*P)
    wall = 45.0 if beams == 128 else 9.0
    fr = [synth.frame(i, beams, cols, wall_radius=wall) if beams == 128 else
          synth.frame(i, beams, cols, wall_radius=wall, ego0=(10.0, 5.0, 1.0), dego=(0.9, 0.5, 0.25)) for i in range(6)]
    dev = [torch.from_numpy(f[0]).cuda() for f in fr]
    for i in range(6):                                     # a warm ring buffer and a previous combined map
        g.Process_pointcloud(dev[i], fr[i][1], fr[i][2])
        g.combine_maps(device_outputs=True)
    torch.cuda.synchronize()
    launches0 = g.stats()["kernel_launches"]
    pc, ego, T = dev[5], fr[5][1], np.ascontiguousarray(fr[5][2], dtype=np.float64)
    e = (C.c_double * 3)(*[float(v) for v in ego])
    gm, lm, nodes = C.c_float(0), C.c_float(0), C.c_int32(0)
    check(g._L.gvom_graph_probe(g._h, pc.data_ptr(), pc.shape[0], 3, GVOM_F64, e, T.ctypes.data, iters, C.byref(gm), C.byref(lm),
                                C.byref(nodes)), "gvom_graph_probe")
    print(json.dumps({"case": name, "iters": iters, "graph_nodes": nodes.value, "graph_us_per_tick": 1e3 * gm.value,
                      "pdl_launch_us_per_tick": 1e3 * lm.value, "graph_speedup": lm.value / gm.value,
                      "note": "device-resident cloud, maps left in HBM, no L2 flush between ticks, launches issued from C"}),
          flush=True)
