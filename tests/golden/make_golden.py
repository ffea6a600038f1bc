#!/usr/bin/env python3
"""Generate golden vectors by EXECUTING the unmodified reference Gvom.

The reference has no tests or fixtures of its own, so the
golden data for the parity tests are dumps of the reference itself replaying the
scenarios in gvom_b200/synth.py:

  * on a B200 through Numba-CUDA (the primary oracle: compiled PTX semantics)
        gpurun -- python tests/golden/make_golden.py --all --out gpurun_out/golden
  * in the CPU container through NUMBA_ENABLE_CUDASIM=1 (tiny grids only)
        NUMBA_ENABLE_CUDASIM=1 python tests/golden/make_golden.py --scenario tiny \
            --out tests/golden --suffix _sim

The reference is imported from /root/reference/scripts or, on the GPU box, from
the git-ignored copy baseline/_ref/gvom.py.  Nothing of it is copied into the
fixtures except its *outputs*.

Everything compact-index dependent is canonicalised through the index map
(compact ids are scheduling dependent in the reference, gvom.py:1238,1031,1059):
arrays are stored sorted by linear voxel id x + y*xy + z*xy*xy.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
for cand in ("/root/reference/scripts", os.path.join(ROOT, "baseline", "_ref")):
    if os.path.exists(os.path.join(cand, "gvom.py")):
        sys.path.insert(0, cand)
        REF_DIR = cand
        break
else:
    raise SystemExit("reference gvom.py not found")

import numba  # noqa: E402
import numba.cuda  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "baseline"))
import ref_shims  # noqa: E402  (external shims, see baseline/ref_shims.py)

CUDASIM = ref_shims.CUDASIM

import gvom as refgvom  # noqa: E402  (the reference)

from gvom_b200 import synth  # noqa: E402


def sha_i(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def host(a):
    return a.copy_to_host() if hasattr(a, "copy_to_host") else np.asarray(a)


def canon_scan(g, full):
    """Canonical dump of the slot written by the last Process_pointcloud."""
    s = g.last_buffer_index
    idx = host(g.index_buffer[s])
    hit = host(g.hit_count_buffer[s])
    tot = host(g.total_count_buffer[s])
    met = host(g.metrics_buffer[s])
    mnh = host(g.min_height_buffer[s])
    org = host(g.origin_buffer[s])
    ids = np.flatnonzero(idx >= 0).astype(np.int32)
    c = idx[ids]
    codes = np.where(idx >= 0, 0, idx).astype(np.int32)
    d = {"origin": org, "n_occ": np.int64(ids.size),
         "codes_sha": sha_i(codes), "codes_sum": np.int64(codes.sum(dtype=np.int64)),
         "ids_sha": sha_i(ids), "hit_sha": sha_i(hit[c]), "total_sha": sha_i(tot[c]),
         "hit_sum": np.int64(hit.sum()), "total_sum": np.int64(tot.sum()),
         "minh_sha": sha_i(mnh[c])}
    if full:
        d.update({"codes": codes, "ids": ids, "hit": hit[c], "total": tot[c],
                  "metrics": met[c], "minh": mnh[c]})
    return d


def canon_combine(g, out, full):
    idx = host(g.combined_index_map)
    ids = np.flatnonzero(idx >= 0).astype(np.int32)
    c = idx[ids]
    codes = np.where(idx >= 0, 0, idx).astype(np.int32)
    hit = host(g.combined_hit_count)[c]
    tot = host(g.combined_total_count)[c]
    mnh = host(g.combined_min_height)[c]
    d = {"out_origin": out[0], "out_pos": out[1], "out_neg": out[2], "out_rough": out[3],
         "out_vis": out[4], "n_occ": np.int64(ids.size),
         "codes_sha": sha_i(codes), "codes_sum": np.int64(codes.sum(dtype=np.int64)),
         "ids_sha": sha_i(ids), "hit_sha": sha_i(hit), "total_sha": sha_i(tot),
         "hit_sum": np.int64(hit.sum(dtype=np.int64)), "total_sum": np.int64(tot.sum(dtype=np.int64)),
         "minh_sha": sha_i(mnh)}
    if full:
        d.update({"codes": codes, "ids": ids, "hit": hit, "total": tot, "minh": mnh,
                  "metrics": host(g.combined_metrics)[c],
                  "eig": host(g.voxels_eigenvalues)[c],
                  "height": host(g.height_map), "inferred": host(g.inferred_height_map),
                  "x_slope": host(g.x_slope_map), "y_slope": host(g.y_slope_map),
                  "guessed": host(g.guessed_height_delta)})
    return d


def canon_debug(g):
    idx = host(g.combined_index_map)
    ids = np.flatnonzero(idx >= 0)
    vox = g.make_debug_voxel_map()
    return {"voxel": vox[idx[ids]], "height": g.make_debug_height_map(),
            "inferred": g.make_debug_inferred_height_map()}


def run(name, outdir, suffix, full_all):
    P, steps = synth.scenario(name)
    g = refgvom.Gvom(*P)
    rec = {}
    meta = {"scenario": name, "params": list(P), "ref_dir": REF_DIR, "numba": numba.__version__,
            "numpy": np.__version__, "cudasim": CUDASIM, "steps": [], "inputs_sha": []}
    if not CUDASIM:
        dev = numba.cuda.get_current_device()
        meta["device"] = dev.name.decode() if isinstance(dev.name, bytes) else str(dev.name)
        meta["cc"] = list(dev.compute_capability)
    small = name in synth.SMALL_SCENARIOS
    last_scan = max(i for i, s in enumerate(steps) if s[0] == "scan")
    last_comb = max(i for i, s in enumerate(steps) if s[0] == "combine")
    t0 = time.time()
    for i, st in enumerate(steps):
        meta["steps"].append(st[0])
        if st[0] == "scan":
            _, pc, ego, T = st
            meta["inputs_sha"].append(synth.sha(pc) + (synth.sha(T) if T is not None else "-"))
            # the reference transforms the device copy only; hand it a private host copy anyway
            g.Process_pointcloud(np.array(pc, copy=True), ego, None if T is None else T.copy())
            d = canon_scan(g, small or full_all or i == last_scan)
        elif st[0] == "combine":
            meta["inputs_sha"].append("")
            out = g.combine_maps()
            d = canon_combine(g, out, small or full_all or i == last_comb)
        else:
            meta["inputs_sha"].append("")
            d = canon_debug(g)
        for k, v in d.items():
            if isinstance(v, str):
                meta.setdefault("sha", {})[f"s{i}_{k}"] = v
            else:
                rec[f"s{i}_{k}"] = np.asarray(v)
        print(f"[{name}] step {i} {st[0]} done ({time.time()-t0:.1f}s)", flush=True)
    meta["seconds"] = time.time() - t0
    os.makedirs(outdir, exist_ok=True)
    path = os.path.join(outdir, f"{name}{suffix}.npz")
    np.savez_compressed(path, meta=np.array(json.dumps(meta)), **rec)
    print(f"wrote {path} ({os.path.getsize(path)/1e6:.2f} MB)", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenario", action="append", default=[])
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "golden"))
    ap.add_argument("--suffix", default="")
    ap.add_argument("--full", action="store_true", help="full arrays at every step")
    a = ap.parse_args()
    names = list(a.scenario)
    if a.all:
        names = list(synth.SMALL_SCENARIOS) + list(synth.FULL_SCENARIOS)
    for n in names:
        run(n, a.out, a.suffix, a.full)


if __name__ == "__main__":
    main()
