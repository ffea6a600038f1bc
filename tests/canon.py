"""Canonical (compact-id independent) dumps and golden comparison helpers.

A "ref-style" object is anything that carries the reference's attribute names
(gvom.py:50-70): the reference itself (numba device arrays), the CPU oracle
(numpy) or the `.refview()` of the CUDA implementation.  Everything is sorted
by linear voxel id x + y*xy + z*xy*xy, because compact ids are scheduling
dependent in the reference (gvom.py:1238,1031,1059).
"""
import hashlib
import json

import numpy as np


def sha_i(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def host(a):
    return a.copy_to_host() if hasattr(a, "copy_to_host") else np.asarray(a)


def canon_scan(g, full=True):
    """Slot written by the last Process_pointcloud (gvom.py:198-208)."""
    s = g.last_buffer_index
    idx = host(g.index_buffer[s])
    hit, tot = host(g.hit_count_buffer[s]), host(g.total_count_buffer[s])
    met, mnh = host(g.metrics_buffer[s]), host(g.min_height_buffer[s])
    ids = np.flatnonzero(idx >= 0).astype(np.int32)
    c = idx[ids]
    codes = np.where(idx >= 0, 0, idx).astype(np.int32)
    d = {"origin": host(g.origin_buffer[s]), "n_occ": np.int64(ids.size),
         "codes_sha": sha_i(codes), "codes_sum": np.int64(codes.sum(dtype=np.int64)),
         "ids_sha": sha_i(ids), "hit_sha": sha_i(hit[c]), "total_sha": sha_i(tot[c]),
         "hit_sum": np.int64(hit[c].sum(dtype=np.int64)), "total_sum": np.int64(tot[c].sum(dtype=np.int64)),
         "minh_sha": sha_i(mnh[c])}
    if full:
        d.update({"codes": codes, "ids": ids, "hit": hit[c], "total": tot[c],
                  "metrics": met[c], "minh": mnh[c]})
    return d


def canon_combine(g, out, full=True):
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
                  "metrics": host(g.combined_metrics)[c], "eig": host(g.voxels_eigenvalues)[c],
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


# ---------------------------------------------------------------------------
# comparison against a golden .npz
# ---------------------------------------------------------------------------

EXACT = {"origin", "n_occ", "codes", "codes_sum", "ids", "hit", "total", "hit_sum", "total_sum", "minh",
         "out_origin", "out_pos", "out_neg", "out_vis"}
# floating-point fields: (rtol, atol).  north_star: 1e-4 relative on heights, slope,
# roughness; moments/eigenvalues are float-atomic sums (order dependent).
FLOAT = {"metrics": (1e-4, 1e-6), "eig": (1e-4, 1e-6), "out_rough": (1e-4, 1e-9),
         "height": (1e-4, 1e-9), "inferred": (1e-4, 1e-9), "x_slope": (1e-4, 1e-9),
         "y_slope": (1e-4, 1e-9), "guessed": (1e-4, 1e-9), "voxel": (1e-4, 1e-6)}
ROW_FIELDS = {"metrics", "eig", "voxel"}          # per-cell rows, subsampled in full-size fixtures


class Golden:
    def __init__(self, path):
        self.z = np.load(path)
        self.meta = json.loads(str(self.z["meta"]))
        self.stride = int(self.meta.get("row_stride", 1))
        self.sha = self.meta.get("sha", {})

    def fields(self, step):
        pre = f"s{step}_"
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}

    def compare(self, step, got, what="", skip=()):
        """Return a list of mismatch descriptions for one step (empty = parity)."""
        bad = []
        gold = self.fields(step)
        kind = self.meta["steps"][step]
        for k, gv in gold.items():
            if k in skip:
                continue
            if k not in got:
                bad.append(f"{what} step {step} ({kind}): field {k} missing from result")
                continue
            v = np.asarray(got[k])
            if self.stride > 1 and (k in ROW_FIELDS or (kind == "debug" and k in ("height", "inferred"))):
                v = v[::self.stride]
            if v.shape != gv.shape:
                bad.append(f"{what} step {step} ({kind}): {k} shape {v.shape} != golden {gv.shape}")
                continue
            if k in EXACT:
                if not np.array_equal(v, gv):
                    n = int(np.sum(v != gv))
                    bad.append(f"{what} step {step} ({kind}): {k} differs in {n}/{gv.size} entries")
            else:
                rtol, atol = FLOAT.get(k, (1e-4, 1e-9))
                v64, g64 = v.astype(np.float64), gv.astype(np.float64)
                ok = np.isclose(v64, g64, rtol=rtol, atol=atol, equal_nan=True)
                if k in ("eig", "voxel") and v64.ndim == 2:
                    cols = slice(0, 3) if k == "eig" else slice(5, 8)
                    ok[:, cols] |= eig_ok(v64, g64, k)
                if not ok.all():
                    j = np.argmax(np.abs(v64 - g64) * ~ok)
                    bad.append(f"{what} step {step} ({kind}): {k} {int((~ok).sum())}/{gv.size} outside "
                               f"rtol={rtol}; worst got {v64.flat[j]!r} golden {g64.flat[j]!r}")
        for k, want in self.sha.items():
            pre = f"s{step}_"
            if k.startswith(pre) and k[len(pre):] in got:
                if got[k[len(pre):]] != want:
                    bad.append(f"{what} step {step} ({kind}): sha256 of {k[len(pre):-4]} differs")
        return bad


def eig_degenerate_gap():
    """Rows whose smaller eigenvalue gap is at most this fraction of the spread (l0 - l2) count as
    near-degenerate: only those may use the wide eigenvalue band (see eig_ok)."""
    return EIG_GAP_GATE


EIG_GAP_GATE = 0.05


def eig_ok(v64, g64, kind="eig"):
    """Eigenvalue tolerance: |error| <= 1e-4 * lambda_max of the row.  The eigenvalues come from the
    closed-form trig solve on FLOAT32 covariances (gvom.py:1437-1487); near a repeated eigenvalue acos() turns a
    1-ulp float32 difference of the input (which the reference's own float-atomic order produces run to run)
    into ~sqrt(ulp) of the angle, so rows that an explicit gap test flags as near-degenerate
    (min(l0-l1, l1-l2) <= EIG_GAP_GATE * (l0-l2) on the golden values) may use 1e-3 * lambda_max.
    Measured (scripts/eigen_error_hist.py, profiles/eigen_hist_*.json): non-degenerate rows stay below 1e-6,
    all rows below 1e-4 on every golden scenario.  kind "voxel": debug rows hold (l0-l1, l1-l2, l2) in columns 5..7.
    Returns a boolean [rows, 3] array."""
    if kind == "eig":
        w, v = g64[:, 0:3], v64[:, 0:3]
        l0, l1, l2 = w[:, 0], w[:, 1], w[:, 2]
    else:
        w, v = g64[:, 5:8], v64[:, 5:8]
        l2 = w[:, 2]; l1 = l2 + w[:, 1]; l0 = l1 + w[:, 0]
    lmax = np.maximum(np.abs(l0), np.abs(l2))
    spread = np.maximum(l0 - l2, 1e-300)
    degenerate = np.minimum(l0 - l1, l1 - l2) <= EIG_GAP_GATE * spread
    err = np.abs(v - w)
    tight = err <= 1e-4 * lmax[:, None] + 1e-7
    wide = degenerate[:, None] & (err <= 1e-3 * lmax[:, None] + 1e-7)
    return tight | wide
