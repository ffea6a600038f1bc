"""World-size-1 MultiGpuGvom loop (partial + finish kernels on one GPU) for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29577")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gvom_b200 import synth  # noqa: E402
from gvom_b200.multi import MultiGpuGvom  # noqa: E402

torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda:0"))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
g = MultiGpuGvom(*synth.params_tuple(), device=0, exchange=sys.argv[2] if len(sys.argv) > 2 else "auto")
fr = [synth.frame(i, 128, 2048) for i in range(4)]
dev = [torch.from_numpy(f[0]).cuda() for f in fr]
torch.cuda.synchronize()
for i in range(steps):
    k = i % 4
    g.Process_pointcloud(dev[k], fr[k][1], fr[k][2])
    g.combine_maps(device_outputs=True)
print("done", g.exchange, g.stats())
dist.destroy_process_group()
