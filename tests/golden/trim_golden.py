#!/usr/bin/env python3
"""Trim the full reference dumps (gpurun_out/golden/*.npz, made by make_golden.py on
the B200) into the committed fixtures tests/golden/*.npz.

Small scenarios are kept whole.  Full-size scenarios keep every integer array
that is compact (ids, hit, total, min height, the 2-D maps, the 5-tuples), drop
the dense 4M-voxel `codes` arrays (their sha256 and sum stay in `meta`), and
keep every STRIDE-th row (voxel-id order) of the per-cell float arrays
(metrics -> float32, eig, debug voxel rows) and of the per-column debug rows.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from gvom_b200 import synth  # noqa: E402

STRIDE = 8


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    for name in synth.SMALL_SCENARIOS + synth.FULL_SCENARIOS:
        p = os.path.join(src, name + ".npz")
        if not os.path.exists(p):
            continue
        z = np.load(p)
        meta = json.loads(str(z["meta"]))
        out = {}
        full = name in synth.FULL_SCENARIOS
        for k in z.files:
            if k == "meta":
                continue
            a = z[k]
            field = k.split("_", 1)[1]
            if full:
                if field == "codes":
                    continue
                if field in ("metrics", "eig", "voxel") or (field in ("height", "inferred") and a.ndim == 2 and a.shape[1] in (3, 7)):
                    a = a[::STRIDE]
                    if a.dtype == np.float64:
                        a = a.astype(np.float32)
            out[k] = a
        meta["row_stride"] = STRIDE if full else 1
        dst = os.path.join(HERE, name + ".npz")
        np.savez_compressed(dst, meta=np.array(json.dumps(meta)), **out)
        print(f"{name}: {os.path.getsize(p)/1e6:.2f} MB -> {os.path.getsize(dst)/1e6:.2f} MB")


if __name__ == "__main__":
    main()
