"""Replay a synth scenario on a ref-style Gvom object and compare every step
with a golden dump of the executed reference."""
import os

import numpy as np

import canon
from gvom_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name, suffix=""):
    return canon.Golden(os.path.join(GOLDEN_DIR, f"{name}{suffix}.npz"))


def replay(make, name, gold=None, view=lambda g: g, what="", full=True, on_step=None, skip=()):
    """make(params) -> object with the Gvom API; view(obj) -> ref-style state.
    Returns (list of mismatches, list of canonical dumps)."""
    P, steps = synth.scenario(name)
    g = make(P)
    bad, dumps = [], []
    for i, st in enumerate(steps):
        if st[0] == "scan":
            _, pc, ego, T = st
            if gold is not None:
                want = gold.meta["inputs_sha"][i]
                have = synth.sha(pc) + (synth.sha(T) if T is not None else "-")
                assert want == have, f"scenario {name} step {i}: synthetic input differs from the golden run's"
            g.Process_pointcloud(np.array(pc, copy=True), ego, None if T is None else T.copy())
            d = canon.canon_scan(view(g), full)
        elif st[0] == "combine":
            out = g.combine_maps()
            d = canon.canon_combine(view(g), out, full)
        else:
            d = canon.canon_debug(view(g))
        dumps.append(d)
        if gold is not None:
            bad += gold.compare(i, d, what, skip)
        if on_step:
            on_step(i, st, d)
    return bad, dumps
