#!/usr/bin/env python3
"""bench.py -- scans/s and per-scan latency of Process_pointcloud + combine_maps.

  python bench.py [--gpus N] [--steps K] [--warmup W]            (N>1: under torchrun)
  python bench.py --impl reference ...                           (reference arm)

One step = one synthetic OS1-128 scan (262,144 points) processed into
a 256x256x64 grid followed by one combine_maps() -- BASELINE.json configs[1] with 4
ring slots; at N>1 every rank owns one sensor stream with 2 ring slots (configs[2],
SURVEY 8d: 2 slots per sensor) and the combine is exchanged across ranks.  Prints ONE
JSON line (rank 0).

  value  : scans/s with the cloud already in HBM and the maps left in HBM
           (CUDA events around every step on the launching stream; L2 flushed
           between steps, outside the timed region)
  e2e    : the same through the public drop-in API with HOST buffers: pinned float64
           cloud in, numpy maps out, H2D and D2H inside the timed region
  roofline / rooflines : per-kernel CUDA-event times (library profiling events on the
           same stream) against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
           and, for the ray-cast kernel, the L2 atomic throughput measured here by the
           library's microbenchmark
  cpu_baseline : the CPU oracle port (oracle/) timed on this box's host cores on a
           bounded sample of the same workload (reported, not a target)

--impl reference: the UNMODIFIED reference class (baseline/_ref/gvom.py or
/root/reference/scripts/gvom.py) through Numba-CUDA on the same GPU with the two
external shims of baseline/ref_shims.py -- the reference has no CPU implementation
other than Numba's simulator (about a day per scan at this size).  If
Numba cannot drive the GPU the arm falls back to the CPU oracle port on all host
threads (cpu_baseline.kind = "port").
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from gvom_b200 import synth  # noqa: E402

METRIC = "scans/sec (Process_pointcloud+combine_maps, OS1-128 262,144 pts, 256x256x64 grid)"
BEAMS, COLS = 128, 2048
NFRAMES = 8

# --config: BASELINE.json configs[1] is the headline (default); configs[3] / configs[4] are the stress regimes
# (ray-cast / atomic bound, and DDA-length + buffer-merge bound).  The stress lines are extra evidence committed under
# profiles/, not the driver's line.
CONFIGS = {
    "os1_128": dict(workload="configs[1]: synthetic OS1-128 scan (128x2048=262,144 pts), 256x256x64 grid @0.4/0.2 m",
                    params={}, points=BEAMS * COLS, frames=NFRAMES, max_points=None),
    "dense": dict(workload="configs[3]: dense stress, 2,097,152-point aggregated cloud (16 x 128x1024 scans) into a 1024x1024x128 grid @0.1 m",
                  params=dict(xy_resolution=0.1, z_resolution=0.1, xy_size=1024, z_size=128, buffer_size=4),
                  points=16 * 128 * 1024, frames=3, max_points=16 * 128 * 1024),
    "long_range": dict(workload="configs[4]: long-range stress, OS1-128 scan with a 200 m wall, 256x256x64 grid @0.4/0.2 m, 16 ring slots",
                       params=dict(buffer_size=16), points=BEAMS * COLS, frames=NFRAMES, max_points=None),
}
CONFIG = "os1_128"


def config_frames(name, rank=0):
    """Frames of a stress configuration (same generators as the parity scenarios of gvom_b200/synth.py)."""
    out = []
    if name == "dense":
        for i in range(CONFIGS[name]["frames"]):
            ego = (100.0 + 0.4 * i, 50.0 + 0.1 * i, 1.0)
            T = synth.pose_matrix(ego, 0.01 * i)
            pc = np.concatenate([synth.synthetic_scan(128, 1024, seed=5000 + 16 * i + k, ego=ego) for k in range(16)], axis=0)
            out.append((np.ascontiguousarray(pc), ego, T))
    elif name == "long_range":
        for i in range(CONFIGS[name]["frames"]):
            out.append(synth.frame(i, BEAMS, COLS, wall_radius=200.0, seed_base=1000 * rank))
    return out


def config_dict(nsens, slots):
    """`config` of the JSON line -- the SAME dict in both arms (the driver compares them)."""
    w = CONFIGS[CONFIG]["workload"] + f", {slots} ring slots per sensor, {nsens} sensor stream(s)"
    if nsens > 1:
        w += f" (configs[2]: {slots * nsens} slots in total, SURVEY 8d: 2 per sensor; one combine_maps over all of them per step)"
    return {"workload": w, "sensors": nsens, "slots_per_sensor": slots, "frames": NFRAMES,
            "l2": "flushed between steps (256 MiB memset on the same stream, outside every timed interval)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def frames(rank=0):
    """NFRAMES consecutive frames of the SURVEY 8(d) stream (sensor `rank` on a 2 m ring)."""
    if CONFIG != "os1_128":
        return config_frames(CONFIG, rank)
    out = []
    for i in range(NFRAMES):
        pc, ego, T = synth.frame(i, BEAMS, COLS, seed_base=1000 * rank)
        if rank:
            a = np.pi / 4 * rank
            T = T.copy()
            T[0, 3] += 2.0 * np.cos(a)
            T[1, 3] += 2.0 * np.sin(a)
        out.append((pc, ego, T))
    return out


def cpu_baseline(sample_steps=4, threads=None):
    """CPU oracle port on the host cores: bounded sample of the same workload."""
    from oracle import gvom_oracle
    L = gvom_oracle.lib()
    cores = threads or os.cpu_count()
    L.gvo_set_threads(cores)
    g = gvom_oracle.OracleGvom(*synth.params_tuple())
    fr = frames()
    g.Process_pointcloud(*fr[0]); g.combine_maps()                       # warm-up (page faults)
    t0 = time.perf_counter()
    for i in range(sample_steps):
        g.Process_pointcloud(*fr[(i + 1) % NFRAMES])
        g.combine_maps()
    dt = time.perf_counter() - t0
    return {"value": sample_steps / dt, "unit": "scans/s", "cores": cores, "kind": "port",
            "sample": f"{sample_steps} scans+combines of the bench workload, oracle/gvom_oracle.c "
                      f"(OpenMP on the ray-cast and merge loops, {cores} threads; moment passes serial)",
            "ms_per_step": 1e3 * dt / sample_steps, "work": g.work}


def cudasim_baseline(timeout_s=150):
    """The north_star's second reported baseline: the UNMODIFIED reference under NUMBA_ENABLE_CUDASIM on this box's
    host cores.  The simulator runs every CUDA thread as a Python thread (about a day per OS1-128 scan),
    so the sample is a reduced configuration: 4 x 24 points into a 16 x 16 x 8 grid, one timed scan + combine."""
    out = os.path.join(ROOT, "gpurun_out", "ref_probe_cudasim_bench.json")
    try:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        env = dict(os.environ, NUMBA_ENABLE_CUDASIM="1")
        t0 = time.perf_counter()
        subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "ref_probe.py"), "--xy", "16", "--z", "8", "--beams", "4",
                        "--cols", "24", "--iters", "1", "--out", out], env=env, check=True, timeout=timeout_s,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        d = json.load(open(out))
        return {"kind": "cudasim-reduced", "value": d["scans_per_sec"], "unit": "scans/s", "cores": os.cpu_count(),
                "ms_per_step": d["end_to_end_ms_p50"], "wall_s": time.perf_counter() - t0,
                "sample": "unmodified reference, NUMBA_ENABLE_CUDASIM=1, 96 points into 16x16x8 (the full workload is "
                          "262,144 points into 256x256x64: ~2700x the points, 2048x the voxels), 1 timed scan + combine"}
    except Exception as ex:
        return {"kind": "cudasim-reduced", "error": repr(ex)}


def run_reference(args):
    """Reference arm: unmodified reference Gvom via Numba-CUDA on this GPU."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = {"impl": "reference", "metric": METRIC, "unit": "scans/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {}}
    # the same workload as our arm: N sensors feed ONE reference Gvom (README.md:49) with `slots` ring slots per
    # sensor; a step is one scan of every sensor + one combine_maps
    nsens = max(1, args.gpus)
    slots = args.slots_per_sensor or (2 if nsens > 1 else 4)
    slots = CONFIGS[CONFIG]["params"].get("buffer_size", slots)
    base["config"] = config_dict(nsens, slots)
    base["arm"] = f"{nsens} sensor stream(s) into ONE unmodified reference Gvom ({slots * nsens} ring slots) on one GPU"
    fr = [frames(r) for r in range(nsens)]
    try:
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        for cand in ("/root/reference/scripts", os.path.join(ROOT, "baseline", "_ref")):
            if os.path.exists(os.path.join(cand, "gvom.py")):
                sys.path.insert(0, cand)
                break
        else:
            raise RuntimeError("reference gvom.py not present (baseline/_ref)")
        if args.ref_mode == "oracle":
            raise RuntimeError("--ref-mode oracle")
        import ref_shims  # noqa: F401
        import numba.cuda
        import gvom as refgvom
        if refgvom.__file__.startswith(os.path.join(ROOT, "gvom_b200")):
            raise RuntimeError("import gvom resolved to the B200 shim, not the reference")
        g = refgvom.Gvom(*synth.params_tuple(**dict(CONFIGS[CONFIG]["params"], buffer_size=slots * nsens)))
        flush = numba.cuda.device_array(256 << 20, dtype=np.uint8)
        ts = []
        for i in range(args.warmup + args.steps):
            numba.cuda.cudadrv.driver.device_memset(flush, 0, 256 << 20)       # L2 flush, outside the timed region
            numba.cuda.synchronize()
            t0 = time.perf_counter()
            for r in range(nsens):
                pc, ego, T = fr[r][i % NFRAMES]
                g.Process_pointcloud(pc, ego, T)
            g.combine_maps()
            numba.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        ts = ts[args.warmup:]
        v = nsens * len(ts) / sum(ts)
        base.update({"value": v, "ms_per_step": 1e3 * sum(ts) / len(ts), "p50_latency_ms": 1e3 * statistics.median(ts),
                     "value_p50": nsens / statistics.median(ts),
                     "cpu_baseline": {"value": v, "unit": "scans/s", "cores": 1, "kind": "reference",
                                      "sample": f"{len(ts)} steps; unmodified reference class through Numba-CUDA "
                                                "(PTX JIT compute_90->sm_100) on the same B200, 1 host thread; "
                                                "shims: baseline/ref_shims.py"},
                     "e2e": {"value": v, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                             "value_p50": nsens / statistics.median(ts)},
                     "gpu_launches": None})
    except Exception as ex:  # Numba cannot drive this GPU: CPU oracle port on all host threads
        steps = max(1, min(args.steps, 6))
        cb = cpu_baseline(sample_steps=steps)
        base.update({"value": cb["value"], "ms_per_step": cb["ms_per_step"], "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "note": f"reference via Numba unavailable ({type(ex).__name__}: {ex}); timed the CPU oracle port"})
    print(json.dumps(base), flush=True)


def multi_parity_check(g_class, P_rank, slots, world, rank, dev, exchange, steps=3):
    """N > 1: a fresh multi-GPU Gvom replays `steps` frames per sensor; every rank also feeds ALL sensors' frames
    through one single-GPU Gvom with world * slots ring slots on its own GPU and compares: the delivered maps
    (ints exact, roughness 1e-4) and the 3-D state of the rows it owns (codes, hit / pass counts, min heights
    exact; moments 1e-4).  -> dict for the JSON line; "ok" is the AND over all ranks."""
    import hashlib
    import torch
    import torch.distributed as dist
    from gvom_b200.gvom import Gvom
    st = torch.cuda.Stream(device=dev)
    g = g_class(*P_rank, device=dev, stream=st.cuda_stream, torch_stream=st, exchange=exchange)
    PN = list(P_rank); PN[4] = slots * world
    ref = Gvom(*PN, device=dev)
    allfr = [frames(r)[:steps] for r in range(world)]
    res = {"maps_equal": True, "codes_equal": True, "counts_equal": True, "moments_close": True}
    sha = hashlib.sha256()
    for s in range(steps):
        g.Process_pointcloud(*allfr[rank][s])
        out = g.combine_maps()
        for r in range(world):
            ref.Process_pointcloud(*allfr[r][s])
        want = ref.combine_maps()
        for a, b in zip(out, want):
            same = np.allclose(a, b, rtol=1e-4, atol=1e-9, equal_nan=True) if a.dtype.kind == "f" else np.array_equal(a, b)
            res["maps_equal"] &= bool(same)
        S, Z = g.xy_size, g.z_size
        v, w = g.refview(), ref.refview()
        rows = slice(None)
        if getattr(g, "_rows", False) and g.exchange == "p2p":      # sharded state: the world rows this rank owns
            oy = int(round(out[0][1] / g.xy_resolution))
            rows = slice((rank - oy) % world, None, world)
        a = v.combined_index_map.reshape(Z, S, S)[:, rows, :]
        b = w.combined_index_map.reshape(Z, S, S)[:, rows, :]
        ca, cb = np.where(a >= 0, 0, a), np.where(b >= 0, 0, b)
        res["codes_equal"] &= bool(np.array_equal(ca, cb))
        ia, ib = a[a >= 0], b[b >= 0]
        if ia.size == ib.size:
            res["counts_equal"] &= bool(np.array_equal(v.combined_hit_count[ia], w.combined_hit_count[ib]) and
                                        np.array_equal(v.combined_total_count[ia], w.combined_total_count[ib]) and
                                        np.array_equal(v.combined_min_height[ia], w.combined_min_height[ib]))
            res["moments_close"] &= bool(np.allclose(v.combined_metrics[ia], w.combined_metrics[ib], rtol=1e-4, atol=2e-6))
        else:
            res["counts_equal"] = res["moments_close"] = False
        sha.update(np.ascontiguousarray(cb).tobytes())
        for m in want[1:]:
            sha.update(np.ascontiguousarray(m if m.dtype.kind != "f" else np.round(m, 6)).tobytes())
    ok = all(res.values())
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=f"cuda:{dev}")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    res.update({"ok": bool(int(t.item())), "steps": steps, "ranks": world, "rank0_reference_sha": sha.hexdigest()[:16],
                "against": f"one single-GPU Gvom with {slots * world} ring slots fed with all {world} sensors' frames, on every rank",
                "state": "row shards (world rows (y + origin_y) mod ranks)" if rows != slice(None) else "replicated"})
    g.close(); ref.close()
    return res


def atomic_peak(L, torch, dev):
    """L2 atomic throughput (G atomics/s): random words of a 16 MiB table, and one word."""
    import ctypes as C
    words = 1 << 22
    table = torch.empty(words, dtype=torch.int32, device=f"cuda:{dev}")
    res = {}
    for name, mode in (("spread", 0), ("same_address", 1)):
        ms, n = C.c_float(0), C.c_int64(0)
        rc = L.gvom_bench_atomics(dev, table.data_ptr(), words, 64, mode, 5, C.byref(ms), C.byref(n))
        if rc == 0 and ms.value > 0:
            res[name] = n.value / (ms.value * 1e-3) / 1e9
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-mode", default="numba", choices=["numba", "oracle"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--slots-per-sensor", type=int, default=0, help="ring slots per sensor / rank (default: 4 at N=1, 2 at N>1)")
    ap.add_argument("--config", default="os1_128", choices=list(CONFIGS), help="workload: BASELINE.json configs[1] (default), [3] dense, [4] long_range")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"], help="multi-GPU combine exchange")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    global CONFIG, NFRAMES
    CONFIG = args.config
    NFRAMES = CONFIGS[CONFIG]["frames"]
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from gvom_b200 import _lib
    from gvom_b200.gvom import Gvom

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    torch.cuda.set_device(local)
    dev = local
    multi = world > 1
    if multi:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{dev}"))

    stream = torch.cuda.Stream(device=dev)
    # ring slots per sensor: 4 for the single-sensor configuration (BASELINE configs[1]); 2 per sensor for the
    # multi-sensor one (configs[2], SURVEY 8d "config 3": B = 16 slots for 8 sensors, the reference README's rule)
    slots = args.slots_per_sensor or (2 if multi else 4)
    cfg = CONFIGS[CONFIG]
    P = synth.params_tuple(**dict(dict(buffer_size=slots), **cfg["params"]))
    slots = P[4]
    extra = {"max_points": cfg["max_points"]} if cfg["max_points"] else {}
    if multi:
        from gvom_b200.multi import MultiGpuGvom
        g = MultiGpuGvom(*P, device=dev, stream=stream.cuda_stream, torch_stream=stream, exchange=args.exchange)
    else:
        g = Gvom(*P, device=dev, stream=stream.cuda_stream, **extra)
    fr = frames(rank)
    pinned = [torch.from_numpy(f[0]).pin_memory() for f in fr]
    on_dev = [p.cuda(dev) for p in pinned]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{dev}")
    torch.cuda.synchronize()

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    # device-resident outputs are left in HBM in stream order (no host wait inside the timed region)
    dev_kw = {"wait": False}

    def timed(device_io, steps, warmup, profile=False):
        """-> (per-step ms list [device events], per-step wall ms, stage-time sums).
        device_io: the steps are enqueued back to back -- [L2 flush][event a][scan + combine][event b] per step, no host
        synchronisation inside the region (device-resident outputs are left in HBM in stream order), one barrier +
        synchronize on both sides; the per-step time is b - a, so the flush is outside every timed interval.
        host I/O (e2e) and the profiling pass: one step at a time, wall clock around both calls."""
        ev, wall, stages, pairs = [], [], {}, []
        if device_io and not profile:
            barrier()
        for i in range(warmup + steps):
            k = i % NFRAMES
            with torch.cuda.stream(stream):
                flush.zero_()                                   # L2 flush, outside the timed region
            if not device_io or profile:
                stream.synchronize()
                if multi:
                    dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record(stream)
            if device_io:
                with torch.cuda.stream(stream):                 # the Gvom's stream is the caller's current stream
                    g.Process_pointcloud(on_dev[k], fr[k][1], fr[k][2])
                    out = g.combine_maps(device_outputs=True, **dev_kw)
            else:
                g.Process_pointcloud(pinned[k], fr[k][1], fr[k][2])
                out = g.combine_maps()
            b.record(stream)
            assert out is not None
            if device_io and not profile:
                if i >= warmup:
                    pairs.append((a, b))
                continue
            b.synchronize()
            t1 = time.perf_counter()
            if i >= warmup:
                ev.append(a.elapsed_time(b))
                wall.append(1e3 * (t1 - t0))
                if profile:
                    for kk, vv in g.stage_times().items():
                        stages[kk] = stages.get(kk, 0.0) + vv
        if pairs:
            barrier()
            ev = [a.elapsed_time(b) for a, b in pairs]
            wall = list(ev)
        return ev, wall, stages

    # ---- timed regions
    barrier()
    sampler = ClockSampler(dev) if rank == 0 else None
    ev_dev, wall_dev, _ = timed(True, args.steps, args.warmup)
    barrier()
    ev_e2e, wall_e2e, _ = timed(False, args.steps, args.warmup)
    barrier()
    clocks = sampler.stop() if sampler else None
    launches0 = g.stats()["kernel_launches"]
    calls0 = g.stats()["process_calls"]
    # ---- per-kernel CUDA-event times (separate pass: the extra event records stay out of the numbers above)
    g.set_profiling(True)
    psteps = min(args.steps, 50)
    _, _, stages = timed(True, psteps, 3, profile=True)
    g.set_profiling(False)
    st = g.stats()
    launches_per_step = (st["kernel_launches"] - launches0) / max(1, st["process_calls"] - calls0)

    def agg(ms_list):
        t = torch.tensor([sum(ms_list)], dtype=torch.float64, device=f"cuda:{dev}")
        if multi:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    tot_dev, tot_e2e = agg(ev_dev), agg(wall_e2e)
    parity = None
    if multi:
        from gvom_b200.multi import MultiGpuGvom
        parity = multi_parity_check(MultiGpuGvom, P, slots, world, rank, dev, args.exchange)
    if rank != 0:
        if multi:
            dist.destroy_process_group()
        return

    value = world * args.steps / (tot_dev * 1e-3)
    e2e = world * args.steps / (tot_e2e * 1e-3)
    hbm, hbm_src = peaks()
    V = P[2] * P[2] * P[3]
    B = P[4]
    N = cfg["points"]
    sources = min(B, args.warmup + args.steps) + 1
    stage_ms = {k: v / psteps for k, v in stages.items()}
    # algorithmic bytes per launch (DESIGN.md "kernels")
    work = None
    atom = atomic_peak(_lib.lib(), torch, dev)
    try:
        from oracle.gvom_oracle import OracleGvom
        if CONFIG == "os1_128":                 # (the stress grids take the CPU oracle minutes: work counts only here)
            o = OracleGvom(*P)
            o.Process_pointcloud(*fr[0])
            work = o.work
    except Exception:
        pass
    # ---- rooflines.  Every stage gets a bound it cannot beat, so frac <= 1:
    #   hbm    algorithmic bytes (SURVEY.md 8d figures where the survey gives one, else the bytes the stage's own data
    #          structures force it to move) / measured copy bandwidth
    #   issue  warp instructions the kernel executed (committed ncu counter capture of this build) / peak issue rate
    #   l2_atomic  atomic sectors the kernel sent to L2 after warp aggregation (same capture) / the atomic issue rate
    #          measured live by gvom_bench_atomics
    # the bound of a stage is the LARGEST of the times these give; frac = that time / the measured stage time.
    S2c = P[2] * P[2]
    counters = {}
    cp_path = os.path.join(ROOT, "profiles", "kernel_counters_r02.json")
    if os.path.exists(cp_path):
        counters = json.load(open(cp_path)).get(CONFIG, {})
    sm_clock_hz = (clocks or {}).get("sm_max_mhz") or 1965.0
    issue_peak = 148 * 4 * sm_clock_hz * 1e6            # warp instructions / s (4 schedulers per SM, 1 per clock)
    cells_scan = (work or {}).get("n_occ") or (work or {}).get("cells") or 30000
    st_now = g.stats()
    cells_comb = max(1, int(st_now.get("combined_cells", 0))) if not multi else None
    abytes = {
        "scan_points": 24 * N,                                            # the cloud as stored (float64 x 3), read once
        "scan_cells": 4 * V + cells_scan * (27 * 4 + 80 + 8),             # the slot map once (group mask) + per cell: look-ups, metrics, counts
        "merge_codes": 4 * V * sources + 4 * V,                           # B slot maps + previous map read, combined map written
        "merge_cells": (cells_comb or 100000) * (4 * sources + 92 + 68),  # per combined cell: look-ups, >= one source record, the merged record
        "maps": 4 * V + 20 * S2c,                                         # one pass over the combined map + the 2-D outputs
    }
    stage_kernels = {"scan_points": ["k_scan_points"], "scan_cells": ["k_scan_cells"], "merge_codes": ["k_merge_rows", "k_merge_codes"],
                     "merge_cells": ["k_merge_cells2"], "maps": ["k_column_maps", "k_surface_maps2"]}
    rooflines = {}
    for k, bts in abytes.items():
        t = stage_ms.get(k, 0) * 1e-3
        if t <= 0 or multi:
            continue
        bounds = {"hbm": bts / (hbm * 1e9)}
        inst = sum(v.get("inst_executed", 0) for kn, v in counters.items() if any(kn.startswith(p) for p in stage_kernels[k]))
        atoms = sum(v.get("atom_sectors", 0) + v.get("red_sectors", 0) for kn, v in counters.items() if any(kn.startswith(p) for p in stage_kernels[k]))
        if inst:
            bounds["issue"] = inst / issue_peak
        if atoms and atom.get("spread"):
            bounds["l2_atomic"] = atoms / (atom["spread"] * 1e9)
        which = max(bounds, key=bounds.get)
        r = {"bound": which, "frac": bounds[which] / t, "ms": stage_ms[k], "bound_ms": {b: 1e3 * v for b, v in bounds.items()},
             "algorithmic_bytes": bts}
        if which == "hbm":
            r.update({"achieved": bts / t / 1e9, "peak": hbm, "unit": "GB/s"})
        elif which == "issue":
            r.update({"achieved": inst / t / 1e9, "peak": issue_peak / 1e9, "unit": "G warp instructions/s", "instructions": inst})
        else:
            r.update({"achieved": atoms / t / 1e9, "peak": atom["spread"], "unit": "G atomic sectors/s", "atomic_sectors": atoms})
        rooflines[k] = r
    if work and stage_ms.get("scan_points", 0) > 0:
        rooflines.setdefault("scan_points", {})["increments_delivered_per_s_G"] = (2 * work["n_in"] + work["dda_steps"]) / (stage_ms["scan_points"] * 1e-3) / 1e9
    # the whole step against the HBM roofline of the path (206.8 MB per scan + combine at configs[1])
    step_bytes = 24 * N + 12 * V + 4 * V * sources + 4 * V + 4 * V + 20 * S2c
    step_roof = {"bound": "hbm", "algorithmic_bytes": step_bytes, "achieved": step_bytes / (tot_dev / args.steps * 1e-3) / 1e9,
                 "peak": hbm, "unit": "GB/s", "frac": step_bytes / (tot_dev / args.steps * 1e-3) / 1e9 / hbm,
                 "note": "SURVEY 8d per-stage bytes summed (cloud, index build, merge, column pass, outputs) / ms_per_step"}
    if multi:
        # N > 1: the per-rank step is scan + push of the scan to the row owners + the combine of the own rows from local
        # mirrors + the 2-D pushes -- reported as stage times, the single-GPU kernel rooflines do not apply
        step_roof["note"] += "; per rank (weak scaling: the same bytes on every GPU, exchange traffic not counted)"
    traffic = {}
    for k, pref in stage_kernels.items():
        vals = [v["dram_read"] + v["dram_write"] for kn, v in counters.items() if any(kn.startswith(p) for p in pref) and "dram_read" in v]
        if vals:
            traffic[k] = sum(vals)
    kernels = {k: v for k, v in stage_ms.items() if k in rooflines}
    dom = max(kernels, key=lambda k: kernels[k], default=None)
    roof = dict(rooflines[dom], kernel=dom, traffic=traffic.get(dom), peak_source=hbm_src,
                traffic_source="profiles/kernel_counters_r02.json (ncu single-pass counters of this build, per launch)") if dom else \
        dict(step_roof, kernel="whole step", traffic=None, peak_source=hbm_src)

    line = {
        "metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(world, slots),
        "io": {"exchange": (getattr(g, "exchange", None) or "") + (": ring slots mirrored to the row owners at scan time, row-sharded combine" if getattr(g, "_mirror", False) and getattr(g, "exchange", "") == "p2p" else " (partial merge + replicated finish)" if world > 1 else ""),
               "value": "cloud resident in HBM (float64 Nx3), maps left in HBM in stream order; steps enqueued back to back, CUDA events around both calls of every step on the launching stream (L2 flush between the intervals), summed",
               "e2e": "pinned host float64 Nx3 cloud in, numpy maps out (pinned), per-step wall clock around both calls"},
        "value_p50": world / (statistics.median(ev_dev) * 1e-3),
        "p50_latency_ms": statistics.median(ev_dev),
        "e2e": {"value": e2e, "unit": "scans/s", "h2d_bytes_per_step": int(fr[0][0].nbytes),
                "d2h_bytes_per_step": 20 * P[2] * P[2], "p50_latency_ms": statistics.median(wall_e2e),
                "value_p50": world / (statistics.median(wall_e2e) * 1e-3),
                "p50_device_ms": statistics.median(ev_e2e)},
        "gpu_launches": int(round(launches_per_step * args.steps)),
        "gpu_launches_per_step": launches_per_step,
        "clocks": clocks,
        "roofline": roof,
        "rooflines": rooflines,
        "step_roofline": step_roof,
        "stage_ms": stage_ms,
        "l2_atomic_peak_gops": atom,
        "work_per_scan": work,
    }
    if world == 1 and CONFIG == "os1_128":
        # the callers either side of the path, same workload, wall clock per tick (p50 of 60):
        # what the reference node hands over (pageable float64 array) and the PointCloud2 payload it received
        # (48-byte records), maps or the node's int8 OccupancyGrid payloads out
        try:
            from gvom_b200.node import PointCloud2Payload
            msgs = [PointCloud2Payload.from_xyz(f[0], 48) for f in fr]
            raw = [m.data.tobytes() for m in msgs]

            def wall(fn, n=60):
                ts = []
                for i in range(n + 8):
                    t0 = time.perf_counter()
                    fn(i % NFRAMES)
                    ts.append(1e3 * (time.perf_counter() - t0))
                return statistics.median(ts[8:])
            line["e2e_variants_p50_ms"] = {
                "pageable_numpy_f64_in__maps_out": wall(lambda k: (g.Process_pointcloud(fr[k][0], fr[k][1], fr[k][2]), g.combine_maps())),
                "pinned_f64_in__int8_grids_out": wall(lambda k: (g.Process_pointcloud(pinned[k], fr[k][1], fr[k][2]), g.combine_maps_grids())),
                "pointcloud2_bytes48_in__maps_out": wall(lambda k: (g.Process_pointcloud2(raw[k], msgs[k].n_points, 48, fr[k][1], fr[k][2]), g.combine_maps())),
                "pointcloud2_bytes48_in__int8_grids_out": wall(lambda k: (g.Process_pointcloud2(raw[k], msgs[k].n_points, 48, fr[k][1], fr[k][2]), g.combine_maps_grids())),
                "host_float64_conversion_the_node_does_first": wall(lambda k: msgs[k].to_xyz_array(), 6),
            }
        except Exception as ex:
            line["e2e_variants_p50_ms"] = {"error": repr(ex)}
    if world == 1 and not args.no_cpu_baseline and CONFIG == "os1_128":
        try:
            line["cpu_baseline"] = cpu_baseline()
        except Exception as ex:
            line["cpu_baseline"] = {"error": repr(ex)}
    if world == 1 and not args.no_cpu_baseline and CONFIG == "os1_128":
        line["cpu_baseline_cudasim"] = cudasim_baseline()
    if parity is not None:
        line["parity_check"] = parity
    print(json.dumps(line), flush=True)
    if multi:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        raise SystemExit("multi-GPU parity check failed: " + json.dumps(parity))


if __name__ == "__main__":
    main()
