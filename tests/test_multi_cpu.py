"""World-size-2 `gloo` test of the multi-GPU combine PROTOCOL on CPU (no CUDA).

The device kernels cannot run here, so each rank models its partial result with the CPU oracle
and numpy: ring slots folded into the common frame as an encoded grid (1<<26 = occupied, else
summed passes), exactly what gvom_combine_partial produces.  The ranks exchange with the same
collectives the NCCL path uses (all_gather of the header, all_reduce(sum) of the grid), decode
like gvom_combine_finish does, and rank 0 checks the result against ONE oracle Gvom that holds
both ranks' scans.  This pins: the encoding, the order independence of the fold, the header
logic (merge_headers) and the previous-map rule applied after the cross-rank sum."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OCC = 1 << 26


def shifted(src, d, S, Z, fill):
    """src indexed [z,y,x] in its own frame -> array in the combined frame (combined voxel (x,y,z) looks at
    source voxel (x+dx, y+dy, z+dz), gvom.py:1017-1027); outside the source: `fill`."""
    out = np.full((Z, S, S), fill, src.dtype)
    dx, dy, dz = (int(v) for v in d)

    def rng(n, dd):
        lo, hi = max(0, -dd), min(n, n - dd)
        return (slice(lo, hi), slice(lo + dd, hi + dd)) if hi > lo else (slice(0, 0), slice(0, 0))
    (zc, zs), (yc, ys), (xc, xs) = rng(Z, dz), rng(S, dy), rng(S, dx)
    out[zc, yc, xc] = src[zs, ys, xs]
    return out


def partial_grid(o, origin, S, Z):
    """numpy model of gvom_combine_partial's code grid for the oracle `o` (its own valid slots)."""
    occ = np.zeros((Z, S, S), bool)
    passes = np.zeros((Z, S, S), np.int64)
    for i in range(o.buffer_size):
        if o.origin_buffer[i] is None:
            continue
        idx = shifted(o.index_buffer[i].reshape(Z, S, S), origin - o.origin_buffer[i], S, Z, -1)
        occ |= idx >= 0
        passes += np.where(idx < -1, -idx - 1, 0)
    return np.where(occ, OCC, np.minimum(passes, OCC - 1)).astype(np.int32)


def finish_codes(total, prev_codes, prev_origin, origin, S, Z):
    """numpy model of the code part of gvom_combine_finish: canonical codes (occupied -> 0)."""
    occ = total >= OCC
    c = -1 - (total.astype(np.int64) & (OCC - 1))
    if prev_codes is not None:
        p = shifted(prev_codes, origin - prev_origin, S, Z, -1)
        take = ~occ & (p >= 0) & (c >= -11)                    # gvom.py:1058
        add = ~occ & (p < -1)
        c = np.where(add, c + p + 1, c)
        occ = occ | take
    return np.where(occ, 0, c).astype(np.int32)


def worker(rank, world, port, result):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gvom_b200 import synth
    from gvom_b200.multi import HEADER_DOUBLES, merge_headers
    from oracle.gvom_oracle import OracleGvom
    from test_multi_gpu import sensor_frames
    S, Z, B = 32, 8, 2
    P1 = synth.params_tuple(xy_size=S, z_size=Z, buffer_size=B, robot_radius=2.0)
    PN = synth.params_tuple(xy_size=S, z_size=Z, buffer_size=B * world, robot_radius=2.0)
    fr = sensor_frames(world, 4, beams=8, cols=64, wall=5.0)
    mine = OracleGvom(*P1)
    ref = OracleGvom(*PN) if rank == 0 else None           # one Gvom holding all ranks' scans
    prev_codes, prev_origin = None, None
    ok = True
    for step in range(4):
        if not (rank == 1 and step == 0):                      # rank 1 has no data at the first combine
            mine.Process_pointcloud(*fr[step][rank])
        have = mine.origin_buffer[mine.last_buffer_index] is not None
        org = mine.origin_buffer[mine.last_buffer_index] if have else np.zeros(3)
        hdr = torch.zeros(HEADER_DOUBLES, dtype=torch.float64)
        hdr[0], hdr[2:5] = float(have), torch.from_numpy(np.asarray(org, np.float64))
        heads = [torch.zeros_like(hdr) for _ in range(world)]
        dist.all_gather(heads, hdr)
        origin, _ = merge_headers(torch.stack(heads).numpy())
        assert origin is not None
        grid = partial_grid(mine, origin, S, Z) if have else np.zeros((Z, S, S), np.int32)
        t = torch.from_numpy(grid.copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        codes = finish_codes(t.numpy(), prev_codes, prev_origin, origin, S, Z)
        # every rank must hold the same result (replicated state)
        both = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(both, torch.from_numpy(codes.copy()))
        ok &= bool(torch.equal(both[0], both[1]))
        prev_codes, prev_origin = codes, origin
        if rank == 0:
            # the big ring (B*world slots) receives this step's scans; with rank 1 silent at step 0 the
            # slots it overwrites are exactly the ones the per-rank rings (B slots each) drop
            for r in range(world):
                if not (r == 1 and step == 0):
                    ref.Process_pointcloud(*fr[step][r])
            ref.combine_maps()
            want = np.where(ref.combined_index_map >= 0, 0, ref.combined_index_map).reshape(Z, S, S)
            ok &= bool(np.array_equal(codes, want))
    result[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_protocol_gloo():
    world = 2
    with mp.Manager() as m:
        result = m.dict()
        mp.spawn(worker, args=(world, 29631, result), nprocs=world, join=True)
        assert dict(result) == {0: True, 1: True}


def test_merge_headers():
    from gvom_b200.multi import merge_headers
    h = np.zeros((3, 8))
    assert merge_headers(h)[0] is None
    h[1, :5] = [1, 7, 10, 20, -3]
    h[2, :5] = [1, 9, 10, 20, -3]
    org, counts = merge_headers(h)
    assert list(org) == [10, 20, -3] and list(counts) == [0, 7, 9]
    h[2, 2] = 11
    with pytest.raises(RuntimeError):
        merge_headers(h)


def fold_slots(slots, origin, S, Z):
    """numpy model of the single merge pass over ring slots (k_merge_rows / gvom.py:1009-1035, order independent):
    slots = [(index_map (V,), origin (3,))] -> (occupied mask, summed passes) in the combined frame."""
    occ = np.zeros((Z, S, S), bool)
    passes = np.zeros((Z, S, S), np.int64)
    for idx, org in slots:
        a = shifted(np.asarray(idx).reshape(Z, S, S), origin - org, S, Z, -1)
        occ |= a >= 0
        passes += np.where(a < -1, -a - 1, 0)
    return occ, passes


@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_plane_sharded_column_exchange_model(nranks):
    """Executable model of the NEXT multi-GPU design: the combined state stays sharded
    by WORLD plane, owner(z) = (z + origin_z) mod N, and the 2-D stage gets the three per-column facts that span
    planes through two small reductions:
      (a) lowest occupied voxel + its min height:  MIN over ranks of the 64-bit key  z << 32 | float32 bits of min_h
      (b) lowest free voxel:                       MIN over ranks of z
      (c) positive-obstacle window sums:           SUM over ranks of (sum hit, sum total) of the cells with hit > 10
    Every rank only looks at its own planes; the reduced facts must reproduce the oracle's height map, inferred
    height map and positive-obstacle map (gvom.py:515-590)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from gvom_b200 import synth
    from oracle.gvom_oracle import OracleGvom
    from test_multi_gpu import sensor_frames
    S, Z = 32, 16
    P = synth.params_tuple(xy_size=S, z_size=Z, buffer_size=4, robot_radius=2.0)
    o = OracleGvom(*P)
    fr = sensor_frames(2, 3, beams=32, cols=512, wall=5.5)
    for step in range(3):
        for r in range(2):
            o.Process_pointcloud(*fr[step][r])
        _, pos_ref, _, _, _ = o.combine_maps()           # two ego steps: the origin (and with it plane ownership) moves
    pos_ref = pos_ref.T                                   # oracle maps are [x, y]; everything below is [y, x]
    z_res, pos_thr, robot_h, slope_thr = P[1], P[6], P[9], P[8]
    oz = int(o.combined_origin[2])
    cmap = o.combined_index_map.reshape(Z, S, S)          # [z, y, x]
    hit, tot, minh = o.combined_hit_count, o.combined_total_count, o.combined_min_height
    INF = np.uint64(0xFFFFFFFFFFFFFFFF)
    key_occ = np.full((nranks, S, S), INF, np.uint64)     # [rank, y, x]
    z_free = np.full((nranks, S, S), 1 << 30, np.int64)
    owner = (np.arange(Z) + oz) % nranks
    for r in range(nranks):
        for z in np.flatnonzero(owner == r):
            plane = cmap[z]
            occ = plane >= 0
            bits = np.zeros((S, S), np.uint64)
            bits[occ] = minh[plane[occ]].view(np.uint32).astype(np.uint64)
            key = (np.uint64(z) << np.uint64(32)) | bits
            key_occ[r] = np.where(occ, np.minimum(key_occ[r], key), key_occ[r])
            z_free[r] = np.where(plane < -1, np.minimum(z_free[r], z), z_free[r])
    k = key_occ.min(axis=0)                               # reduction (a)
    zf = z_free.min(axis=0)                               # reduction (b)
    has = k != INF
    zo = (k >> np.uint64(32)).astype(np.int64)
    mh = (k & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.float32).astype(np.float64)
    h_model = np.where(has, (zo + mh + oz) * z_res, -1000.0)           # [y, x]
    h_ref = o.height_map.T                                # oracle maps are [x, y]
    outside_disc = ~((h_ref > -1000.0) & ~has)            # the ego disc's assumed ground is not a column fact
    assert np.array_equal(h_model[has], h_ref[has])
    assert (h_ref[outside_disc & ~has] == -1000.0).all()
    inf_model = np.where(zf < (1 << 30), (oz + zf) * z_res, -1000.0)
    assert np.array_equal(inf_model, o.inferred_height_map.T)
    # (c): the window depends on the (now global) height; each rank sums inside it over ITS planes only
    h_all = h_ref                                         # incl. the disc default, as the 2-D stage sees it
    lo = np.floor((h_all + pos_thr) / z_res - oz).astype(np.int64) + 1
    hi = np.floor((h_all + robot_h) / z_res - oz).astype(np.int64)
    ok = (lo >= 0) & (lo < Z) & (hi >= 0) & (hi < Z)
    sums = np.zeros((nranks, 2, S, S), np.int64)
    for r in range(nranks):
        for z in np.flatnonzero(owner == r):
            plane = cmap[z]
            inwin = ok & (lo <= z) & (z <= hi) & (plane >= 0)
            big = np.zeros((S, S), bool)
            big[inwin] = hit[plane[inwin]] > 10
            sums[r, 0][big] += hit[plane[big]]
            sums[r, 1][big] += tot[plane[big]]
    sh, st = sums[:, 0].sum(axis=0).astype(np.float64), sums[:, 1].sum(axis=0).astype(np.float64)     # reduction (c)
    dens = np.where(st > 0, sh / np.where(st > 0, st, 1.0), sh)
    pos_model = np.where(ok, (dens * 100.0).astype(np.int32), 0)
    steep = ~(np.sqrt(o.x_slope_map ** 2 + o.y_slope_map ** 2) < slope_thr).T
    pos_model = np.where(steep, 100, pos_model)
    assert np.array_equal(pos_model, pos_ref)
    # the scenario exercises both paths: steep cells (100) and density values from the window sums
    assert has.sum() > 20 and (pos_ref == 100).sum() > 0 and ((pos_ref > 0) & ~steep).sum() > 0


# ---------------------------------------------------------------------------------------------------------------
# Mirrored ring slots (gvom_mirror.cuh): executable model of the push protocol of k_push_scan and of what a rank may
# read.  numpy only; the CUDA kernels are checked against a single-GPU Gvom in tests/test_multi_gpu.py.
# ---------------------------------------------------------------------------------------------------------------
def test_newest_origin_host_logic():
    from gvom_b200.multi import MIRROR_ENTRY_INTS, newest_origin
    t = np.zeros((3, 2, MIRROR_ENTRY_INTS), np.int32)
    assert newest_origin(t) == (None, None)
    ego = np.array([12.5, -3.25, 1.0])
    t[1, 0, :5] = [4, 10, 20, -3, 77]
    t[1, 1, :5] = [5, 11, 21, -3, 80]                       # the newer scan of rank 1
    t[1, 1, 8:14] = ego.view(np.int32)
    t[2, 0, :5] = [9, 99, 99, 99, 1]                        # a later rank is not consulted once one is found
    org, e = newest_origin(t)
    assert list(org) == [11, 21, -3] and np.array_equal(e, ego)


@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_mirror_push_protocol_model(nranks):
    """One ring slot of one rank, overwritten by a sequence of scans whose origin moves (rows change owner): after
    every push, every rank's mirror must equal the scan on the rows that rank owns under the scan's origin --
    the only rows a combine reads -- although nothing is ever cleaned up at ranks that lost a row."""
    rng = np.random.default_rng(7)
    S, Z, spr = 256, 4, 1
    nseg = S * Z * spr
    mirrors = [np.full((nseg, 256), -1, np.int32) for _ in range(nranks)]       # per rank: map segments of the mirror
    masks = [np.zeros(nseg, np.uint32) for _ in range(nranks)]
    held = np.zeros((nranks, nseg), np.uint32)                                  # pusher's memory of what each rank holds
    seg_row = (np.arange(nseg) // spr) % S
    oy = 0
    for step in range(12):
        oy += int(rng.integers(-2, 3))                                          # the ego moves: ownership shifts
        scan = np.full((nseg, 256), -1, np.int32)
        known = rng.random(nseg) < 0.4
        for s in np.flatnonzero(known):
            g = rng.integers(0, 32, 3)                                          # a few known 8-voxel groups
            for gi in g:
                scan[s, 8 * gi:8 * gi + 8] = rng.integers(-50, 40, 8)
        word = np.zeros(nseg, np.uint32)
        for gi in range(32):
            word |= ((scan[:, 8 * gi:8 * gi + 8] != -1).any(axis=1).astype(np.uint32) << np.uint32(gi))
        own = (seg_row + oy) % nranks
        # ---- k_push_scan: only the owner of a row is written; wipe where it holds older codes and the scan knows nothing
        for s in range(nseg):
            r = own[s]
            if word[s] == 0 and held[r, s] == 0:
                continue
            mirrors[r][s] = scan[s]
            masks[r][s] = word[s]
            held[r, s] = word[s]
        # ---- what a combine reads: the rows each rank owns under THIS origin
        for r in range(nranks):
            mine = own == r
            assert np.array_equal(mirrors[r][mine], scan[mine]), f"step {step} rank {r}: codes"
            assert np.array_equal(masks[r][mine], word[mine]), f"step {step} rank {r}: masks"
        # every segment is owned by exactly one rank
        assert np.array_equal(np.bincount(own, minlength=nranks).sum(), nseg)
