"""A few steps of the bench workload (device-resident cloud) for ncu. No timing printed."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gvom_b200 import Gvom, synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
host_io = len(sys.argv) > 2 and sys.argv[2] == "host"
g = Gvom(*synth.params_tuple())
fr = [synth.frame(i, 128, 2048) for i in range(4)]
pin = [torch.from_numpy(f[0]).pin_memory() for f in fr]
dev = [p.cuda() for p in pin]
torch.cuda.synchronize()
for i in range(steps):
    k = i % 4
    g.Process_pointcloud(pin[k] if host_io else dev[k], fr[k][1], fr[k][2])
    g.combine_maps(device_outputs=not host_io)
print("done", g.stats())
