#!/usr/bin/env python3
"""A/B timing of the kernel builds kept behind GVOM_VARIANT (gvom_api.cu VAR_*) on the bench workload:
   python scripts/variant_times.py [steps]
For every mask: CUDA-event time of the whole step (device-resident cloud and maps, L2 flushed between
steps, like bench.py's `value`) and the per-kernel stage times of the library's profiling events.
Also times the host-facing variants of a tick: numpy f64 cloud / PointCloud2 payload in, maps / int8 grids out."""
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gvom_b200 import Gvom, synth  # noqa: E402
from gvom_b200.node import PointCloud2Payload  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
NAMES = {0: "current", 1: "old surface", 2: "old merge", 4: "old gather", 8: "old cells", 16: "rows NB6", 32: "gather2 2 blocks/SM", 64: "DMA outputs", 15: "all old"}
stream = torch.cuda.Stream()
g = Gvom(*synth.params_tuple(), stream=stream.cuda_stream)
fr = [synth.frame(i, 128, 2048) for i in range(8)]
pin = [torch.from_numpy(f[0]).pin_memory() for f in fr]
dev = [p.cuda() for p in pin]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()


def run(n, profile=False):
    ev, st = [], {}
    for i in range(n):
        k = i % 8
        with torch.cuda.stream(stream):
            flush.zero_()
        stream.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        g.Process_pointcloud(dev[k], fr[k][1], fr[k][2])
        g.combine_maps(device_outputs=True)
        b.record(stream)
        b.synchronize()
        ev.append(a.elapsed_time(b))
        if profile:
            for kk, vv in g.stage_times().items():
                st[kk] = st.get(kk, 0.0) + vv / n
    return ev, st


res = {}
for mask in (0, 64, 0):
    g.set_variant(mask)
    run(12)
    ev, _ = run(steps)
    g.set_profiling(True)
    _, st = run(20, profile=True)
    g.set_profiling(False)
    keep = ("raycast", "index", "moments", "gather", "merge_codes", "merge_cells", "maps")
    line = {"mask": mask, "name": NAMES[mask], "step_us_mean": 1e3 * statistics.mean(ev), "step_us_p50": 1e3 * statistics.median(ev),
            "stage_us": {k: round(1e3 * st[k], 2) for k in keep}}
    res[f"{mask}:{NAMES[mask]}"] = line
    print(json.dumps(line), flush=True)
g.set_variant(0)

# host-facing ticks (wall clock around the calls, p50 of `steps`)
msgs = [PointCloud2Payload.from_xyz(f[0], 48) for f in fr]           # Ouster-like 48-byte records
raw = [m.data.tobytes() for m in msgs]
rawpin = [torch.from_numpy(m.data.copy()).pin_memory() for m in msgs]


def wall(fn, n):
    ts = []
    for i in range(n + 8):
        t0 = time.perf_counter()
        fn(i % 8)
        ts.append(1e3 * (time.perf_counter() - t0))
    return statistics.median(ts[8:])


host = {
    "numpy_f64_in__maps_out": wall(lambda k: (g.Process_pointcloud(fr[k][0], fr[k][1], fr[k][2]), g.combine_maps()), steps),
    "pinned_f64_in__maps_out": wall(lambda k: (g.Process_pointcloud(pin[k], fr[k][1], fr[k][2]), g.combine_maps()), steps),
    "pinned_f64_in__grids_out": wall(lambda k: (g.Process_pointcloud(pin[k], fr[k][1], fr[k][2]), g.combine_maps_grids()), steps),
    "pc2_bytes48_in__maps_out": wall(lambda k: (g.Process_pointcloud2(raw[k], msgs[k].n_points, 48, fr[k][1], fr[k][2]), g.combine_maps()), steps),
    "pc2_bytes48_in__grids_out": wall(lambda k: (g.Process_pointcloud2(raw[k], msgs[k].n_points, 48, fr[k][1], fr[k][2]), g.combine_maps_grids()), steps),
    "pc2_pinned48_in__grids_out": wall(lambda k: (g.Process_pointcloud2(rawpin[k], msgs[k].n_points, 48, fr[k][1], fr[k][2]), g.combine_maps_grids()), steps),
    "ros_numpy_like_f64_conversion_only": wall(lambda k: msgs[k].to_xyz_array(), 10),
}


def pipelined(n):
    """async combine: scan i+1 is enqueued before the maps of combine i are awaited"""
    pend = None
    t0 = time.perf_counter()
    for i in range(n):
        k = i % 8
        g.Process_pointcloud(pin[k], fr[k][1], fr[k][2])
        nxt = g.combine_maps_async()
        if pend is not None:
            pend.result()
        pend = nxt
    pend.result()
    return 1e3 * (time.perf_counter() - t0) / n


g.set_variant(64)
host["pinned_f64_in__maps_out__DMA_outputs"] = wall(lambda k: (g.Process_pointcloud(pin[k], fr[k][1], fr[k][2]), g.combine_maps()), steps)
host["pinned_f64_in__grids_out__DMA_outputs"] = wall(lambda k: (g.Process_pointcloud(pin[k], fr[k][1], fr[k][2]), g.combine_maps_grids()), steps)
g.set_variant(0)
pipelined(10)
host["pinned_f64_in__maps_out__async_pipelined_ms_per_tick"] = pipelined(steps)
print(json.dumps({"host_tick_p50_ms": host}), flush=True)
res["host_tick_p50_ms"] = host
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/variant_times.json", "w"), indent=1)
