#!/usr/bin/env python3
"""Small probes behind statements in DESIGN.md: (1) whole-step time of a trivial workload (the per-kernel launch /
dependency floor of the 8-kernel pipeline), (2) zero-copy vs chunked-DMA input for a pinned float64 cloud."""
import json, os, statistics, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gvom_b200 import Gvom, synth

if len(sys.argv) > 1 and sys.argv[1] == "h2d":
    g = Gvom(*synth.params_tuple())
    fr = [synth.frame(i, 128, 2048) for i in range(8)]
    pin = [torch.from_numpy(f[0]).pin_memory() for f in fr]
    ts = []
    for i in range(100):
        k = i % 8
        t0 = time.perf_counter()
        g.Process_pointcloud(pin[k], fr[k][1], fr[k][2]); g.combine_maps()
        ts.append(1e3 * (time.perf_counter() - t0))
    print(json.dumps({"GVOM_H2D": os.environ.get("GVOM_H2D", "zero-copy"), "pinned_f64_tick_p50_ms": statistics.median(ts[20:])}))
    sys.exit(0)

stream = torch.cuda.Stream()
out = {}
for name, P, beams, cols, wall in (("tiny 64x64x16, 4k pts", synth.params_tuple(xy_size=64, z_size=16, robot_radius=2.0), 16, 256, 9.0),
                                   ("OS1-128 262k pts, 256x256x64", synth.params_tuple(), 128, 2048, 45.0)):
    g = Gvom(*P, stream=stream.cuda_stream)
    fr = [synth.frame(i, beams, cols, wall_radius=wall) for i in range(4)]
    dev = [torch.from_numpy(f[0]).cuda() for f in fr]
    ev = []
    for i in range(60):
        k = i % 4
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream.synchronize()
        a.record(stream)
        g.Process_pointcloud(dev[k], fr[k][1], fr[k][2]); g.combine_maps(device_outputs=True)
        b.record(stream); b.synchronize()
        ev.append(1e3 * a.elapsed_time(b))
    out[name] = {"step_us_p50_no_l2_flush": statistics.median(ev[10:]), "kernels_per_step": 8}
print(json.dumps(out))
for mode in ("zero-copy", "dma"):
    env = dict(os.environ); env["GVOM_H2D"] = mode
    print(subprocess.run([sys.executable, __file__, "h2d"], env=env, capture_output=True, text=True).stdout.strip())
