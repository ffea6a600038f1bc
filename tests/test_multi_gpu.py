"""Multi-GPU combine: every exchange must equal one Gvom that holds every rank's ring slots.

 * test_mirrored_rows_equals_single (1 GPU): several handles on the same device play the ranks of the default
   combine -- mirrored ring slots (gvom_mirror_attach: every scan pushed to the owners of its world rows) + row-sharded
   combine (gvom_combine_finish_rows); all blocks in plain device memory.
 * test_partial_finish_equals_single (1 GPU): the generic exchange (gvom_combine_partial / gvom_combine_finish), peer
   buffers read directly or one summed grid + gathered records (what NCCL delivers).
 * test_nccl_two_ranks (>= 2 GPUs): MultiGpuGvom under torchrun -- NCCL, p2p generic, p2p mirrored, late-joining rank.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import canon
from gvom_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sensor_frames(nranks, nsteps, beams=16, cols=256, wall=9.0):
    """frames[step][rank] = (cloud, ego, T): common ego per step, sensors on a ring."""
    out = []
    for i in range(nsteps):
        row = []
        for r in range(nranks):
            pc, ego, T = synth.frame(i, beams, cols, wall_radius=wall, seed_base=1000 * r, ego0=(10.0, 5.0, 1.0),
                                     dego=(0.9, 0.5, 0.25))
            T = T.copy()
            T[0, 3] += 0.5 * np.cos(2.0 * r)
            T[1, 3] += 0.5 * np.sin(2.0 * r)
            row.append((pc, ego, T))
        out.append(row)
    return out


def local_exchange(ranks, reduced=False):
    """Emulate the exchange for handles living in one process.  reduced=False: every "rank" reads
    every other rank's grid / records directly (what the peer-to-peer exchange does over NVLink);
    reduced=True: one summed grid + gathered records (what the NCCL exchange delivers)."""
    import torch
    from gvom_b200._lib import GVOM_HOST, GVOM_NO_DATA, RECORD_FLOATS, check
    g0 = ranks[0]
    L, V = g0._L, g0.voxel_count
    dev = f"cuda:{g0.device}"
    org = (C.c_double * 3)()
    grids, masks, recs, counts = [], [], [], []
    origin = None
    for g in ranks:
        if L.gvom_newest_origin(g._h, org) != GVOM_NO_DATA and origin is None:
            origin = [org[0], org[1], org[2]]
    assert origin is not None
    o = (C.c_double * 3)(*origin)
    cap = int(min(V, g0.buffer_size * g0.max_points))
    for g in ranks:
        grid = torch.zeros(V, dtype=torch.int32, device=dev)
        msk = torch.zeros(V // 256 + 2, dtype=torch.int32, device=dev)
        rec = torch.empty((cap, RECORD_FLOATS), dtype=torch.float32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        check(L.gvom_combine_partial(g._h, o, grid.data_ptr(), msk.data_ptr(), rec.data_ptr(), cap, cnt.data_ptr(), None, 0, 0, None),
              "partial")
        torch.cuda.synchronize()
        grids.append(grid); masks.append(msk); recs.append(rec); counts.append(cnt)
    n = len(ranks)
    parr = lambda ps: (C.c_void_p * len(ps))(*[C.c_void_p(int(p)) for p in ps])
    if reduced:
        total = torch.stack(grids).sum(0).to(torch.int32)
        gptrs, mptrs, ng = parr([total.data_ptr()]), None, 1
    else:
        gptrs, mptrs, ng = parr([t.data_ptr() for t in grids]), parr([t.data_ptr() for t in masks]), n
    rptrs, cptrs = parr([t.data_ptr() for t in recs]), parr([t.data_ptr() for t in counts])
    outs = []
    for g in ranks:
        pos, neg, rough, vis = g._out_arrays()
        oo = (C.c_double * 3)()
        check(L.gvom_combine_finish(g._h, o, gptrs, mptrs, ng, rptrs, cptrs, n, cap, None, 0, oo, pos.ctypes.data, neg.ctypes.data,
                                    rough.ctypes.data, vis.ctypes.data, GVOM_HOST, None), "finish")
        outs.append((np.array(list(oo)), pos, neg, rough, vis))
    return outs


def attach_mirrors(ranks):
    """Mirrored ring slots for handles living in one process: one block per "rank" in plain device memory."""
    import torch
    from gvom_b200._lib import check
    L, n = ranks[0]._L, len(ranks)
    nb = C.c_uint64(0)
    check(L.gvom_mirror_block_size(ranks[0]._h, n, C.byref(nb)), "mirror block size")
    blocks = [torch.empty(nb.value, dtype=torch.uint8, device=f"cuda:{ranks[0].device}") for _ in ranks]   # (attach initialises them)
    ptrs = (C.c_void_p * n)(*[C.c_void_p(b.data_ptr()) for b in blocks])
    for r, g in enumerate(ranks):
        check(L.gvom_mirror_attach(g._h, r, n, ptrs), "mirror attach")
    return blocks


def local_exchange_mirror(ranks, epoch, blocks=None):
    """In-process emulation of the MIRRORED row-sharded combine: the scans were pushed by Process_pointcloud; every
    "rank" publishes its epoch flag (phase 16, all ranks), merges its rows from the local mirrors (phase 1 | 32), then
    the 2-D phases as in local_exchange_rows."""
    import torch
    from gvom_b200._lib import GVOM_HOST, GVOM_NO_DATA, GvomRowsLinks, check
    g0 = ranks[0]
    L = g0._L
    dev = f"cuda:{g0.device}"
    n = len(ranks)
    org = (C.c_double * 3)()
    origin = None
    for g in ranks:
        if L.gvom_newest_origin(g._h, org) != GVOM_NO_DATA and origin is None:
            origin = [org[0], org[1], org[2]]
    o = (C.c_double * 3)(*origin)
    nb = C.c_uint64(0)
    check(L.gvom_rows_block_size(g0._h, C.byref(nb)), "block size")
    B = [dict(fh=torch.zeros(64, dtype=torch.int32, device=dev), fr=torch.zeros(64, dtype=torch.int32, device=dev),
              b2d=torch.zeros(nb.value, dtype=torch.uint8, device=dev)) for _ in ranks]
    links = []
    for r in range(n):
        K = GvomRowsLinks()
        K.rank, K.nranks = r, n
        for k, b in enumerate(B):
            K.blocks2d[k] = b["b2d"].data_ptr()
            K.heights_slots[k] = b["fh"].data_ptr() + 4 * r
            K.results_slots[k] = b["fr"].data_ptr() + 4 * r
        K.heights_flags, K.results_flags = B[r]["fh"].data_ptr(), B[r]["fr"].data_ptr()
        links.append(K)
    if blocks is not None:                                  # ranks that have not scanned yet adopt the vehicle position
        from gvom_b200.multi import MIRROR_ENTRY_INTS, newest_origin
        torch.cuda.synchronize()
        nt = n * g0.buffer_size * MIRROR_ENTRY_INTS
        for r, g in enumerate(ranks):
            if L.gvom_newest_origin(g._h, org) == GVOM_NO_DATA:
                table = blocks[r][256:256 + 4 * nt].view(torch.int32).cpu().numpy().reshape(n, g0.buffer_size, MIRROR_ENTRY_INTS)
                o2, ego = newest_origin(table)
                assert o2 is not None and list(o2) == origin
                check(L.gvom_adopt_ego(g._h, (C.c_double * 3)(*[float(v) for v in ego])), "adopt ego")
    outs = []
    # (every flag is published for all ranks before any rank waits for it: 16 epoch, 64 heights, 128 results; + 32 / 256 = the
    # waiting kernels do not publish again)
    for phase in (16, 1 | 32, 64, 2 | 256, 128, 4 | 256):
        for r, g in enumerate(ranks):
            pos, neg, rough, vis = g._out_arrays()
            oo = (C.c_double * 3)()
            check(L.gvom_combine_finish_rows(g._h, o, C.byref(links[r]), epoch, phase, oo, pos.ctypes.data, neg.ctypes.data,
                                             rough.ctypes.data, vis.ctypes.data, GVOM_HOST, None), "rows (mirrored)")
            torch.cuda.synchronize()
            if phase & 4:
                outs.append((np.array(list(oo)), pos, neg, rough, vis))
    return outs, B


def assemble_rows_state(ranks, outs, xy_res):
    """Row-sharded state -> one canonical dump: rank r contributes the world rows (y + origin_y) % n == r."""
    n = len(ranks)
    S, Z = ranks[0].xy_size, ranks[0].z_size
    oy = int(round(outs[0][0][1] / xy_res))
    idx_all = np.full((Z, S, S), -1, np.int32)
    vals = {}
    for r, g in enumerate(ranks):
        v = g.refview()
        y0 = (r - oy) % n
        idx = v.combined_index_map.reshape(Z, S, S)
        own = np.zeros((Z, S, S), bool)
        own[:, y0::n, :] = True
        idx_all[own] = idx[own]
        occ = np.flatnonzero((idx >= 0) & own)
        ids = idx.reshape(-1)[occ]
        for vox, i in zip(occ.tolist(), ids.tolist()):
            vals[vox] = (v.combined_hit_count[i], v.combined_total_count[i], v.combined_min_height[i], v.combined_metrics[i],
                         v.voxels_eigenvalues[i])
    flat = idx_all.reshape(-1)
    ids = np.flatnonzero(flat >= 0).astype(np.int32)
    codes = np.where(flat >= 0, 0, flat).astype(np.int32)
    hit = np.array([vals[i][0] for i in ids.tolist()], np.int32)
    tot = np.array([vals[i][1] for i in ids.tolist()], np.int32)
    mnh = np.array([vals[i][2] for i in ids.tolist()], np.float32)
    met = np.array([vals[i][3] for i in ids.tolist()], np.float32).reshape(-1, 10)
    return {"codes": codes, "ids": ids, "hit": hit, "total": tot, "minh": mnh, "metrics": met}


@pytest.mark.parametrize("nranks,late", [(1, 0), (2, 0), (3, 0), (3, 2), (4, 0), (5, 1)])
def test_mirrored_rows_equals_single(nranks, late):
    """gvom_mirror_attach + gvom_combine_finish_rows (the default multi-GPU combine): every scan is pushed to the owners
    of its rows by Process_pointcloud; the maps every rank delivers and the 3-D state assembled from the ranks' row
    shards equal one Gvom holding every rank's ring slots.  The ego moves 1.25 rows per step, so rows change owner
    between scans.  late: the last rank only starts scanning at that step.  From 4 ranks up (>= 8 ring slots) the row
    merge hands source masks to the cell kernel, which then fetches the records of a cell four at a time."""
    from gvom_b200 import Gvom
    Bs = 2
    kw = dict(xy_size=256, z_size=16, robot_radius=2.0)
    P1, PN = synth.params_tuple(buffer_size=Bs, **kw), synth.params_tuple(buffer_size=Bs * nranks, **kw)
    fr = sensor_frames(nranks, 5, wall=30.0)
    ranks = [Gvom(*P1) for _ in range(nranks)]
    blocks = attach_mirrors(ranks)
    active = lambda step, r: not (late and r == nranks - 1 and step < late)
    for step in range(5):
        for r in range(nranks):
            if active(step, r):
                ranks[r].Process_pointcloud(*fr[step][r])
        outs, _keep = local_exchange_mirror(ranks, step + 1, blocks)
        ref = Gvom(*PN)
        for s2 in range(step + 1):
            for q in range(max(0, s2 - Bs + 1), s2 + 1):
                for r in range(nranks):          # a rank that has not started yet: an empty scan (knows nothing) keeps the ring aligned
                    pc, ego, T = fr[q][r]
                    ref.Process_pointcloud(pc if active(q, r) else np.zeros((0, 3)), ego, T)
            last = ref.combine_maps()
        want = canon.canon_combine(ref.refview(), last)
        for r in range(nranks):
            for a, b, name in zip(outs[r], last, ("origin", "pos", "neg", "rough", "vis")):
                ok = np.allclose(a, b, rtol=1e-4, atol=1e-9, equal_nan=True) if a.dtype.kind == "f" else np.array_equal(a, b)
                assert ok, f"step {step} rank {r}: {name}"
        got = assemble_rows_state(ranks, outs, P1[0])
        for k in ("codes", "ids", "hit", "total", "minh"):
            assert np.array_equal(got[k], want[k]), f"step {step}: {k}"
        assert np.allclose(got["metrics"], want["metrics"], rtol=1e-4, atol=2e-6), f"step {step}: metrics"


def compare_state(a, b, what):
    """a, b: canon_combine dumps. ints exact, floats to tolerance."""
    for k in ("out_origin", "out_pos", "out_neg", "out_vis", "codes", "ids", "hit", "total", "minh"):
        assert np.array_equal(a[k], b[k]), f"{what}: {k}"
    for k in ("out_rough", "height", "inferred", "x_slope", "y_slope", "guessed"):
        assert np.allclose(a[k], b[k], rtol=1e-4, atol=1e-9), f"{what}: {k}"
    assert np.allclose(a["metrics"], b["metrics"], rtol=1e-4, atol=2e-6), f"{what}: metrics"


@pytest.mark.parametrize("nranks,reduced", [(1, False), (2, False), (3, False), (2, True)])
def test_partial_finish_equals_single(nranks, reduced):
    from gvom_b200 import Gvom
    B = 2
    P1 = synth.params_tuple(xy_size=64, z_size=16, buffer_size=B, robot_radius=2.0)
    PN = synth.params_tuple(xy_size=64, z_size=16, buffer_size=B * nranks, robot_radius=2.0)
    fr = sensor_frames(nranks, 5)
    ranks = [Gvom(*P1) for _ in range(nranks)]
    for step in range(5):
        for r in range(nranks):
            ranks[r].Process_pointcloud(*fr[step][r])
        outs = local_exchange(ranks, reduced)
        # One Gvom with B*nranks slots holding the same scans: replay the whole history from
        # scratch (it needs the same chain of "previous combined map" states), feeding before
        # every combine the scans the rank rings hold at that step, newest step last.
        ref = Gvom(*PN)
        for s2 in range(step + 1):
            for q in range(max(0, s2 - B + 1), s2 + 1):
                for r in range(nranks):
                    ref.Process_pointcloud(*fr[q][r])
            last = ref.combine_maps()
        want = canon.canon_combine(ref.refview(), last)
        for r in range(nranks):
            got = canon.canon_combine(ranks[r].refview(), outs[r])
            compare_state(got, want, f"step {step} rank {r}")


def test_nccl_two_ranks(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    script = os.path.join(ROOT, "tests", "multi_rank_check.py")
    for port, exchange, extra in ((29533, "nccl", []), (29534, "p2p", []), (29535, "p2p", ["late"]),
                                  (29536, "p2p", ["grid256"]), (29537, "p2p", ["grid256", "late"])):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                            "--master-addr", "127.0.0.1", "--master-port", str(port), script, exchange] + extra,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        assert "MULTI_RANK_OK" in r.stdout, r.stdout[-2000:]
        print(r.stdout[-300:])
