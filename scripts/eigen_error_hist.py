#!/usr/bin/env python3
"""Error histogram of the covariance eigenvalues (gvom.py:1423-1487) against the golden dumps of the
executed reference: how many rows need more than the surveyed 1e-4 * lambda_max band, and are those
exactly the near-degenerate rows (repeated eigenvalues, where acos() amplifies a 1-ulp float32
difference of the covariance to ~sqrt(ulp) of the angle)?

  python scripts/eigen_error_hist.py --impl oracle|cuda [--out profiles/eigen_hist_rNN.json]

The degeneracy measure is the one tests/canon.py gates its wide band on: the smaller eigenvalue
gap relative to the spread, gap = min(l0-l1, l1-l2) / (l0-l2) (0 = repeated eigenvalue).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import canon  # noqa: E402
import replay  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="oracle", choices=["oracle", "cuda"])
    ap.add_argument("--scenarios", default="small_moving,small_quirks,small_eigen2,os1_64,os1_128,long_range")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    if args.impl == "oracle":
        from oracle.gvom_oracle import OracleGvom
        make, view = (lambda P: OracleGvom(*P)), (lambda g: g)
    else:
        from gvom_b200 import Gvom
        make, view = (lambda P: Gvom(*P)), (lambda g: g.refview())
    edges = [0, 1e-7, 1e-6, 1e-5, 1e-4, 3e-4, 1e-3, 1e-2, 1.0]
    report = {"impl": args.impl, "bins_rel_to_lambda_max": edges, "scenarios": {}}
    for name in args.scenarios.split(","):
        gold = replay.golden(name)
        rows = {"err": [], "gap": [], "lmax": []}

        def on_step(i, st, d, gold=gold, rows=rows):
            if st[0] != "combine":
                return
            g = gold.fields(i)
            if "eig" not in g:
                return
            v = np.asarray(d["eig"], np.float64)[::gold.stride]
            w = g["eig"].astype(np.float64)
            lmax = np.abs(w).max(axis=1)
            spread = np.maximum(w[:, 0] - w[:, 2], 1e-300)
            gap = np.minimum(w[:, 0] - w[:, 1], w[:, 1] - w[:, 2]) / spread
            rows["err"].append(np.abs(v - w).max(axis=1)); rows["gap"].append(gap); rows["lmax"].append(lmax)

        replay.replay(make, name, None, view=view, on_step=on_step)
        err = np.concatenate(rows["err"]); gap = np.concatenate(rows["gap"]); lmax = np.concatenate(rows["lmax"])
        rel = err / np.maximum(lmax, 1e-30)
        rel = np.where(lmax < 1e-12, 0.0, rel)                         # single-point cells: all-zero covariance
        hist, _ = np.histogram(rel, bins=edges)
        over = rel > 1e-4
        gate = canon.eig_degenerate_gap()
        rep = {"rows": int(rel.size), "hist": hist.tolist(), "max_rel": float(rel.max()) if rel.size else 0.0,
               "rows_over_1e-4": int(over.sum()),
               "rows_over_1e-4_not_degenerate": int((over & (gap > gate)).sum()),
               "degenerate_rows(gap<=gate)": int((gap <= gate).sum()), "gate": gate,
               "max_gap_of_rows_over_1e-4": float(gap[over].max()) if over.any() else None,
               "max_rel_of_non_degenerate_rows": float(rel[gap > gate].max()) if (gap > gate).any() else 0.0}
        report["scenarios"][name] = rep
        print(name, json.dumps(rep), flush=True)
    if args.out:
        json.dump(report, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
