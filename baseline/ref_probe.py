#!/usr/bin/env python3
"""Baseline probe: run the UNMODIFIED reference Gvom (Numba CUDA) and time it.

This is a measurement harness for the reference, not part of the new
implementation.  It answers three survey questions:
  1. does the reference's Numba build run on this GPU at all (numba's in-tree
     CUDA target only knows compute capabilities up to 9.0)?
  2. what are its per-call latencies at BASELINE.json's OS1-64 / OS1-128 configs?
  3. what are the algorithmic work counts (occupied voxels, ray steps) of the
     synthetic scan, which the roofline model in SURVEY.md section 8(d) needs?

Usage:
  python baseline/ref_probe.py [--beams 128] [--cols 2048] [--iters 20]
  NUMBA_ENABLE_CUDASIM=1 python baseline/ref_probe.py --xy 32 --z 16 --beams 8 --cols 32 --iters 1
Writes gpurun_out/ref_probe_<beams>x<cols>.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for cand in ("/root/reference/scripts", os.path.join(HERE, "_ref")):
    if os.path.exists(os.path.join(cand, "gvom.py")):
        sys.path.insert(0, cand)
        REF_DIR = cand
        break
else:
    raise SystemExit("reference gvom.py not found (copy it to baseline/_ref/)")

import numba  # noqa: E402
import numba.cuda  # noqa: E402

sys.path.insert(0, HERE)
import ref_shims  # noqa: E402,F401  (external shims, see baseline/ref_shims.py)

import gvom  # noqa: E402


def synthetic_scan(beams, cols, seed, ego, wall_radius=45.0, lidar_height=1.0):
    """OS1-like scan in the SENSOR frame, float64 (N,3), every beam returns.

    Elevation: `beams` angles uniform in [-22.5, +22.5] deg.  Azimuth: `cols`
    uniform in [0, 2pi).  Downward beams hit rolling ground
    z = 0.3 sin(0.15 X) cos(0.1 Y) (world X,Y; 2 fixed-point iterations), capped
    by a cylinder wall of radius `wall_radius`; upward beams hit the wall.
    Gaussian range noise sigma = 0.02 m.
    """
    rng = np.random.default_rng(seed)
    el = np.deg2rad(np.linspace(-22.5, 22.5, beams))[:, None]
    az = (np.arange(cols) * (2 * np.pi / cols))[None, :]
    ce, se = np.cos(el), np.sin(el)
    r_wall = wall_radius / ce
    r = np.broadcast_to(r_wall, (beams, cols)).copy()
    down = np.broadcast_to(se < 0, (beams, cols))
    rg = np.where(down, lidar_height / np.maximum(-se, 1e-9), 1e6)
    for _ in range(2):
        X = ego[0] + rg * ce * np.cos(az)
        Y = ego[1] + rg * ce * np.sin(az)
        zg = 0.3 * np.sin(0.15 * X) * np.cos(0.1 * Y)
        rg = np.where(down, (lidar_height - zg) / np.maximum(-se, 1e-9), 1e6)
    r = np.minimum(r, rg)
    r = r + rng.normal(0.0, 0.02, size=r.shape)
    pts = np.stack([r * ce * np.cos(az), r * ce * np.sin(az), r * se], axis=-1)
    return np.ascontiguousarray(pts.reshape(-1, 3), dtype=np.float64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--beams", type=int, default=128)
    ap.add_argument("--cols", type=int, default=2048)
    ap.add_argument("--xy", type=int, default=256)
    ap.add_argument("--z", type=int, default=64)
    ap.add_argument("--xy-res", type=float, default=0.4)
    ap.add_argument("--z-res", type=float, default=0.2)
    ap.add_argument("--buffer", type=int, default=4)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()

    info = {"ref_dir": REF_DIR, "numba": numba.__version__, "numpy": np.__version__,
            "cudasim": os.environ.get("NUMBA_ENABLE_CUDASIM") == "1",
            "host_cores": os.cpu_count(), "args": vars(a)}
    if not info["cudasim"]:
        dev = numba.cuda.get_current_device()
        info["device"] = dev.name.decode() if isinstance(dev.name, bytes) else str(dev.name)
        info["cc"] = list(dev.compute_capability)
    print(json.dumps(info), flush=True)

    g = gvom.Gvom(a.xy_res, a.z_res, a.xy, a.z, a.buffer, 1.0, 0.5, 0.5, 0.3, 2.0, 4.0, 1.0, 1, 1)

    t_proc, t_comb = [], []
    out = None
    for it in range(a.iters + 2):  # first two iterations are JIT / warm-up
        ego = (100.0 + 0.4 * it, 50.0 + 0.1 * it, 1.0)
        yaw = 0.01 * it
        T = np.eye(4)
        T[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
        T[:3, 3] = ego
        pc = synthetic_scan(a.beams, a.cols, seed=it, ego=ego)
        t0 = time.perf_counter()
        g.Process_pointcloud(pc, ego, T)
        if not info["cudasim"]:
            numba.cuda.synchronize()
        t1 = time.perf_counter()
        out = g.combine_maps()
        t2 = time.perf_counter()
        if it >= 2 or a.iters == 0:
            t_proc.append(t1 - t0)
            t_comb.append(t2 - t1)
        print(f"iter {it}: process {1e3*(t1-t0):.2f} ms  combine {1e3*(t2-t1):.2f} ms", flush=True)

    # work counts of the last processed scan (buffer slot written last)
    slot = g.last_buffer_index
    idx = g.index_buffer[slot].copy_to_host()
    hit = g.hit_count_buffer[slot].copy_to_host()
    tot = g.total_count_buffer[slot].copy_to_host()
    free_passes = int((-idx[idx < -1] - 1).sum())
    res = dict(info)
    res.update({
        "points": int(a.beams * a.cols),
        "voxels": int(a.xy * a.xy * a.z),
        "occupied_voxels_last_scan": int((idx >= 0).sum()),
        "free_voxels_last_scan": int((idx < -1).sum()),
        "hits_in_grid_last_scan": int(hit.sum()),
        "ray_steps_last_scan": int(tot.sum() - hit.sum() + free_passes),
        "combined_cells": int(g.combined_cell_count_cpu),
        "process_ms": [1e3 * t for t in t_proc],
        "combine_ms": [1e3 * t for t in t_comb],
        "process_ms_p50": float(np.median(t_proc)) * 1e3 if t_proc else None,
        "combine_ms_p50": float(np.median(t_comb)) * 1e3 if t_comb else None,
        "out_summary": {
            "origin": [float(v) for v in out[0]],
            "pos_nonzero": int((out[1] > 0).sum()), "pos_eq_100": int((out[1] == 100).sum()),
            "neg_nonzero": int((out[2] > 0).sum()),
            "rough_valid": int((out[3] != -1).sum()),
            "visible": int(out[4].sum()),
            "dtypes": [str(o.dtype) for o in out], "shapes": [list(o.shape) for o in out],
        },
    })
    if res["process_ms_p50"]:
        tot_ms = res["process_ms_p50"] + res["combine_ms_p50"]
        res["end_to_end_ms_p50"] = tot_ms
        res["scans_per_sec"] = 1e3 / tot_ms
    os.makedirs("gpurun_out", exist_ok=True)
    path = a.out or f"gpurun_out/ref_probe_{a.beams}x{a.cols}{'_sim' if info['cudasim'] else ''}.json"
    with open(path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps({k: v for k, v in res.items() if k not in ("process_ms", "combine_ms")}, indent=1))


if __name__ == "__main__":
    main()
