#!/usr/bin/env python3
"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
   * 64x64x16 grid (generic merge kernel), ring wrap, origin shifts, float32 + PointCloud2 + host inputs
   * 256x256x16 grid: the row-segment merge kernels (register build and, GVOM_VARIANT=4, the bulk-copy pipeline)
   * capacity growth, state save / restore, OccupancyGrid kernel, debug exports
Every result is compared with the CPU oracle, so a sanitizer run is also a parity run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import canon  # noqa: E402
from gvom_b200 import Gvom, synth  # noqa: E402
from gvom_b200.node import PointCloud2Payload  # noqa: E402
from oracle.gvom_oracle import OracleGvom  # noqa: E402


def run(xy, wall, variant):
    os.environ["GVOM_VARIANT"] = str(variant)
    P = synth.params_tuple(xy_size=xy, z_size=16, buffer_size=3, robot_radius=2.0)
    g, o = Gvom(*P, device=0, max_points=2048), OracleGvom(*P)
    for i in range(5):
        pc, ego, T = synth.frame(i, 16 if i < 3 else 24, 256, wall_radius=wall, ego0=(10.0, 5.0, 1.0), dego=(0.9, 0.5, 0.25))
        if i == 1:
            msg = PointCloud2Payload.from_xyz(pc, 32)
            g.Process_pointcloud2(msg.data, msg.n_points, 32, ego, T)
        elif i == 2:
            g.Process_pointcloud(pc.astype(np.float32), ego, T)
            pc = pc.astype(np.float32)
        else:
            g.Process_pointcloud(pc, ego, T)          # i == 3 grows the capacity (6144 points > 4096)
        o.Process_pointcloud(pc, ego, T)
        got, want = g.combine_maps(), o.combine_maps()
        for a, b, name in zip(got, want, ("origin", "positive", "negative", "roughness", "visibility")):
            ok = np.allclose(a, b, rtol=1e-4, atol=1e-9) if a.dtype.kind == "f" else np.array_equal(a, b)
            assert ok, (xy, variant, i, name)
        cg, co = canon.canon_combine(g.refview(), got, full=False), canon.canon_combine(o, want, full=False)
        for k in ("codes_sha", "ids_sha", "hit_sha", "total_sha", "minh_sha"):
            assert cg[k] == co[k], (xy, variant, i, k)
        if i == 2:
            blob = g.save_state()
            g.load_state(blob)
    g.occupancy_grids(50, -10, 0)
    g.make_debug_voxel_map(); g.make_debug_height_map(); g.make_debug_inferred_height_map()
    g.close()
    print(f"case xy={xy} variant={variant}: ok", flush=True)


if __name__ == "__main__":
    run(64, 9.0, 0)
    run(256, 30.0, 0)
    run(256, 30.0, 4)
    print("SANITIZER_CASES_OK")
