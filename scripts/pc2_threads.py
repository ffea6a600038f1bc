#!/usr/bin/env python3
"""PointCloud2 payload (pageable bytes, 48-byte records) in -> int8 grids out: tick time against the number of
staging threads (GVOM_COPY_THREADS)."""
import json, os, statistics, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1 and sys.argv[1] == "run":
    import torch
    from gvom_b200 import Gvom, synth
    from gvom_b200.node import PointCloud2Payload
    g = Gvom(*synth.params_tuple())
    fr = [synth.frame(i, 128, 2048) for i in range(8)]
    msgs = [PointCloud2Payload.from_xyz(f[0], 48) for f in fr]
    raw = [m.data.tobytes() for m in msgs]
    ts = []
    for i in range(120):
        k = i % 8
        t0 = time.perf_counter()
        g.Process_pointcloud2(raw[k], msgs[k].n_points, 48, fr[k][1], fr[k][2])
        g.combine_maps_grids()
        ts.append(1e3 * (time.perf_counter() - t0))
    print(json.dumps({"threads": os.environ.get("GVOM_COPY_THREADS", "default"), "tick_p50_ms": statistics.median(ts[20:]),
                      "stage_copy_host_ms": g.stage_times()["stage_copy_host"], "cpus": os.cpu_count()}))
    sys.exit(0)
for n in ("default", "4", "8", "12", "16"):
    env = dict(os.environ)
    if n != "default":
        env["GVOM_COPY_THREADS"] = n
    print(subprocess.run([sys.executable, __file__, "run"], env=env, capture_output=True, text=True).stdout.strip(), flush=True)
