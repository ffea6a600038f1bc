"""Run under torchrun (any world size): MultiGpuGvom over NCCL must equal one Gvom
holding all ranks' slots.  argv[1] = exchange ("nccl", "p2p" or "auto")."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import canon  # noqa: E402
from gvom_b200 import synth  # noqa: E402
from gvom_b200.gvom import Gvom  # noqa: E402
from gvom_b200.multi import MultiGpuGvom  # noqa: E402
from test_multi_gpu import compare_state, sensor_frames  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    B = 2
    xy = 256 if "grid256" in sys.argv else 64                 # the direct exchange needs xy_size % 256 == 0
    P1 = synth.params_tuple(xy_size=xy, z_size=16, buffer_size=B, robot_radius=2.0)
    PN = synth.params_tuple(xy_size=xy, z_size=16, buffer_size=B * world, robot_radius=2.0)
    fr = sensor_frames(world, 4, beams=16, cols=512, wall=30.0) if xy == 256 else sensor_frames(world, 4)
    g = MultiGpuGvom(*P1, device=local, exchange=sys.argv[1] if len(sys.argv) > 1 else "auto",
                     sharded=True if "sharded" in sys.argv else "auto")
    if "late" in sys.argv[2:]:                             # start-up path: rank 1 joins one combine late
        first = MultiGpuGvom(*P1, device=local, exchange=sys.argv[1])
        if rank == 0:
            first.Process_pointcloud(*fr[0][0])
        o = first.combine_maps()
        assert o is not None and o[1].shape == (xy, xy)
        torch.cuda.synchronize()
        dist.barrier()
    for step in range(4):
        g.Process_pointcloud(*fr[step][rank])
        out = g.combine_maps()
        ref = Gvom(*PN, device=local)
        for s2 in range(step + 1):
            for q in range(max(0, s2 - B + 1), s2 + 1):
                for r in range(world):
                    ref.Process_pointcloud(*fr[q][r])
            last = ref.combine_maps()
        compare_state(canon.canon_combine(g.refview(), out), canon.canon_combine(ref.refview(), last),
                      f"step {step} rank {rank}")
    assert sys.argv[1] == "auto" or g.exchange == sys.argv[1], (g.exchange, getattr(g, "_p2p_error", ""))
    dist.barrier()
    if rank == 0:
        print("MULTI_RANK_OK exchange=" + g.exchange + (" sharded" if getattr(g, "_sharded", False) and g.exchange == "p2p" else "") + (" p2p_error=" + getattr(g, "_p2p_error", "") if g.exchange != "p2p" else ""))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
