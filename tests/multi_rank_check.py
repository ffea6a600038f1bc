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


def compare_rows(g, out, ref, last, rank, world, what):
    """Row-sharded state: every rank delivers the full maps; of the 3-D state it holds the world rows it owns."""
    for a, b, name in zip(out, last, ("origin", "pos", "neg", "rough", "vis")):
        ok = np.allclose(a, b, rtol=1e-4, atol=1e-9, equal_nan=True) if a.dtype.kind == "f" else np.array_equal(a, b)
        assert ok, f"{what}: {name}"
    S, Z = g.xy_size, g.z_size
    oy = int(round(out[0][1] / g.xy_resolution))
    y0 = (rank - oy) % world
    v, w = g.refview(), ref.refview()
    a = v.combined_index_map.reshape(Z, S, S)[:, y0::world, :]
    b = w.combined_index_map.reshape(Z, S, S)[:, y0::world, :]
    assert np.array_equal(np.where(a >= 0, 0, a), np.where(b >= 0, 0, b)), f"{what}: codes of the own rows"
    ia, ib = a[a >= 0], b[b >= 0]
    assert np.array_equal(v.combined_hit_count[ia], w.combined_hit_count[ib]), f"{what}: hit"
    assert np.array_equal(v.combined_total_count[ia], w.combined_total_count[ib]), f"{what}: total"
    assert np.array_equal(v.combined_min_height[ia], w.combined_min_height[ib]), f"{what}: min height"
    assert np.allclose(v.combined_metrics[ia], w.combined_metrics[ib], rtol=1e-4, atol=2e-6), f"{what}: metrics"


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    B = 2
    xy = 256 if "grid256" in sys.argv else 64                 # the mirrored, row-sharded combine needs xy_size % 256 == 0
    P1 = synth.params_tuple(xy_size=xy, z_size=16, buffer_size=B, robot_radius=2.0)
    PN = synth.params_tuple(xy_size=xy, z_size=16, buffer_size=B * world, robot_radius=2.0)
    fr = sensor_frames(world, 4, beams=16, cols=512, wall=30.0) if xy == 256 else sensor_frames(world, 4)
    g = MultiGpuGvom(*P1, device=local, exchange=sys.argv[1] if len(sys.argv) > 1 else "auto")
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
        if getattr(g, "_rows", False) and g.exchange == "p2p":
            compare_rows(g, out, ref, last, rank, world, f"step {step} rank {rank}")
        else:
            compare_state(canon.canon_combine(g.refview(), out), canon.canon_combine(ref.refview(), last),
                          f"step {step} rank {rank}")
    assert sys.argv[1] == "auto" or g.exchange == sys.argv[1], (g.exchange, getattr(g, "_p2p_error", ""))
    dist.barrier()
    if rank == 0:
        print("MULTI_RANK_OK exchange=" + g.exchange + (" mirrored rows" if getattr(g, "_rows", False) and g.exchange == "p2p" else "") + (" p2p_error=" + getattr(g, "_p2p_error", "") if g.exchange != "p2p" else ""))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
