"""A few steps of a bench workload (device-resident cloud) for ncu. No timing printed.
   python scripts/profile_step.py [steps] [host] [--config os1_128|dense|long_range]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from gvom_b200 import Gvom, synth  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
steps = int(args[0]) if args else 10
host_io = len(args) > 1 and args[1] == "host"
config = sys.argv[sys.argv.index("--config") + 1] if "--config" in sys.argv else "os1_128"
cfg = bench.CONFIGS[config]
bench.CONFIG, bench.NFRAMES = config, cfg["frames"]
P = synth.params_tuple(**cfg["params"])
g = Gvom(*P, **({"max_points": cfg["max_points"]} if cfg["max_points"] else {}))
fr = bench.frames()[:4]
pin = [torch.from_numpy(f[0]).pin_memory() for f in fr]
dev = [p.cuda() for p in pin]
torch.cuda.synchronize()
for i in range(steps):
    k = i % len(fr)
    g.Process_pointcloud(pin[k] if host_io else dev[k], fr[k][1], fr[k][2])
    g.combine_maps(device_outputs=not host_io)
torch.cuda.synchronize()
print("done", g.stats())
