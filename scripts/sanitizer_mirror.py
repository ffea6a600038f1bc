#!/usr/bin/env python3
"""compute-sanitizer case for the multi-GPU kernels (gvom_mirror.cuh): two "ranks" played by one process on one GPU --
push of every scan to the row owners (k_push_scan, and GVOM_VARIANT=32: its bulk-copy build), flag exchange + source
list (k_mirror_args), own-row merge, cells + heights, bit maps, surface stage, delivery -- with a moving ego (rows
change owner) and a late rank.  Every result is compared with one single-GPU Gvom holding both ranks' scans."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import canon  # noqa: E402
from gvom_b200 import Gvom, synth  # noqa: E402
from test_multi_gpu import assemble_rows_state, attach_mirrors, local_exchange_mirror, sensor_frames  # noqa: E402


def dump_sources(ranks, blocks, origin_m, P1, voxels, n, B):
    """debug: what every mirror of the owner rank holds at the given combined voxels"""
    import torch
    g0 = ranks[0]
    S, Z, V, cap = g0.xy_size, g0.z_size, g0.voxel_count, min(g0.max_points, g0.voxel_count)
    up = lambda v: (v + 255) & ~255
    nsegp = (V // 256 + 2 + 63) & ~63
    o_table = 256
    o_args = up(o_table + n * B * 16 * 4)
    o_held = up(o_args + 65 * 64 + 8)
    o_mir = up(o_held + B * n * nsegp * 4)
    f_gmask = up(V * 4); f_hit = up(f_gmask + nsegp * 4); f_tot = up(f_hit + cap * 4); f_minh = up(f_tot + cap * 4)
    f_met = up(f_minh + cap * 4); mbytes = up(f_met + cap * 80)
    org = [int(round(origin_m[0] / P1[0])), int(round(origin_m[1] / P1[0])), int(round(origin_m[2] / P1[1]))]
    for v in voxels:
        x, y, z = v % S, (v // S) % S, v // (S * S)
        r = (y + org[1]) % n
        blk = blocks[r]
        table = blk[o_table:o_table + n * B * 64].view(torch.int32).cpu().numpy().reshape(n, B, 16)
        for g in range(n):
            for k in range(B):
                e = table[g, k]
                if e[0] == 0:
                    continue
                xs, ys, zs = x + org[0] - e[1], y + org[1] - e[2], z + org[2] - e[3]
                if not (0 <= xs < S and 0 <= ys < S and 0 <= zs < Z):
                    print("  voxel", (x, y, z), "owner", r, "source", (g, k), "outside"); continue
                m0 = o_mir + (g * B + k) * mbytes
                vs = xs + (ys + zs * S) * S
                code = int(blk[m0 + 4 * vs:m0 + 4 * vs + 4].view(torch.int32).item())
                gm = int(blk[m0 + f_gmask + 4 * (vs >> 8):m0 + f_gmask + 4 * (vs >> 8) + 4].view(torch.int32).item())
                rec = None
                if code >= 0:
                    rec = (int(blk[m0 + f_hit + 4 * code:m0 + f_hit + 4 * code + 4].view(torch.int32).item()),
                           int(blk[m0 + f_tot + 4 * code:m0 + f_tot + 4 * code + 4].view(torch.int32).item()),
                           float(blk[m0 + f_minh + 4 * code:m0 + f_minh + 4 * code + 4].view(torch.float32).item()),
                           float(blk[m0 + f_met + 80 * code + 72:m0 + f_met + 80 * code + 80].view(torch.float64).item()))
                print("  voxel", (x, y, z), "owner", r, "source", (g, k), "seq", int(e[0]), "src voxel", (xs, ys, zs), "row owner under slot origin",
                      (ys + int(e[2])) % n, "code", code, "mask word %08x" % (gm & 0xffffffff), "record (hit,tot,minh,n)", rec, flush=True)


def run(variant, nranks=2, steps=3):
    os.environ["GVOM_VARIANT"] = str(variant)
    Bs = 2
    kw = dict(xy_size=256, z_size=8, robot_radius=2.0)
    P1, PN = synth.params_tuple(buffer_size=Bs, **kw), synth.params_tuple(buffer_size=Bs * nranks, **kw)
    fr = sensor_frames(nranks, steps, beams=8, cols=128, wall=20.0)
    ranks = [Gvom(*P1, max_points=4096) for _ in range(nranks)]
    blocks = attach_mirrors(ranks)
    active = lambda step, r: not (r == nranks - 1 and step == 0)          # the last rank joins one combine late
    from oracle.gvom_oracle import OracleGvom
    oracles = [OracleGvom(*P1) for _ in range(nranks)]
    for step in range(steps):
        for r in range(nranks):
            if active(step, r):
                ranks[r].Process_pointcloud(*fr[step][r])
                oracles[r].Process_pointcloud(*fr[step][r])
                a, b = canon.canon_scan(ranks[r].refview()), canon.canon_scan(oracles[r])      # the slot itself, bit for bit
                for k in ("codes", "ids", "hit", "total", "minh"):
                    assert np.array_equal(a[k], b[k]), ("slot", variant, step, r, k, a[k][a[k] != b[k]][:6], b[k][a[k] != b[k]][:6])
        outs, _keep = local_exchange_mirror(ranks, step + 1, blocks)
        ref = Gvom(*PN, max_points=4096)
        for s2 in range(step + 1):
            for q in range(max(0, s2 - Bs + 1), s2 + 1):
                for r in range(nranks):
                    pc, ego, T = fr[q][r]
                    ref.Process_pointcloud(pc if active(q, r) else np.zeros((0, 3)), ego, T)
            last = ref.combine_maps()
        want = canon.canon_combine(ref.refview(), last)
        for r in range(nranks):
            for a, b, name in zip(outs[r], last, ("origin", "pos", "neg", "rough", "vis")):
                ok = np.allclose(a, b, rtol=1e-4, atol=1e-9, equal_nan=True) if a.dtype.kind == "f" else np.array_equal(a, b)
                assert ok, (variant, step, r, name)
        got = assemble_rows_state(ranks, outs, P1[0])
        for k in ("codes", "ids", "hit", "total", "minh"):
            if not np.array_equal(got[k], want[k]):
                bad = np.flatnonzero(got[k] != want[k])
                S = 256
                vox = got["ids"][bad[:8]]
                print("MISMATCH", variant, step, k, len(bad), "of", len(want[k]), "voxels (x,y,z):",
                      [(int(v % S), int((v // S) % S), int(v // (S * S))) for v in vox], "got", got[k][bad[:8]], "want", want[k][bad[:8]],
                      "origin", outs[0][0], flush=True)
                dump_sources(ranks, blocks, outs[0][0], P1, [int(v) for v in vox], nranks, Bs)
                raise AssertionError((variant, step, k))
    print(f"mirror case variant={variant}: ok", flush=True)


if __name__ == "__main__":
    run(0)
    run(32)
    print("SANITIZER_MIRROR_OK")
