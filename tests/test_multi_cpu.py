"""World-size-2 `gloo` tests of the multi-GPU combine PROTOCOLS on CPU (no CUDA).

The device kernels cannot run here, so each rank models its part with the CPU oracle and numpy:
 * test_mirrored_row_sharded_protocol_gloo -- the default combine (mirrored ring slots, state sharded by world rows):
   which rows of which slot a rank may read, the merge of the own rows + own previous rows while the origin moves, the
   assembly of the result.
 * test_two_rank_protocol_gloo -- the generic exchange: ring slots folded into the common frame as an encoded grid
   (1<<26 = occupied, else summed passes), exactly what gvom_combine_partial produces; the collectives of the NCCL path
   (all_gather of the header, all_reduce(sum) of the grid); decoding like gvom_combine_finish.
 * test_mirror_push_protocol_model, test_newest_origin_host_logic, test_merge_headers -- host logic / push bookkeeping.
Rank 0 checks every result against ONE oracle Gvom that holds both ranks' scans."""
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


def mirror_worker(rank, world, port, result):
    """World-size-2 gloo model of the DEFAULT multi-GPU combine (mirrored ring slots, row-sharded state): every rank keeps
    of every rank's slots only the world rows it owns (the rest is filled with garbage: it must never be read), merges its
    own rows + its own rows of the previous combined map, and the assembled result must equal one oracle Gvom holding
    all ranks' scans -- while the ego moves, i.e. the local index of a world row changes from combine to combine."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gvom_b200 import synth
    from oracle.gvom_oracle import OracleGvom
    from test_multi_gpu import sensor_frames
    S, Z, B = 32, 8, 2
    P1 = synth.params_tuple(xy_size=S, z_size=Z, buffer_size=B, robot_radius=2.0)
    PN = synth.params_tuple(xy_size=S, z_size=Z, buffer_size=B * world, robot_radius=2.0)
    fr = sensor_frames(world, 5, beams=8, cols=64, wall=5.0)
    mine = OracleGvom(*P1)
    ref = OracleGvom(*PN) if rank == 0 else None
    rng = np.random.default_rng(100 + rank)
    prev_codes, prev_origin = None, None                       # this rank's shard: only its own rows are meaningful
    ok = True
    ys = np.arange(S)
    for step in range(5):
        mine.Process_pointcloud(*fr[step][rank])
        origin = np.asarray(mine.origin_buffer[mine.last_buffer_index], np.float64)
        # "push": every rank's valid slots travel (gloo stands in for the NVLink stores) ...
        local = [(np.asarray(mine.index_buffer[i]).reshape(Z, S, S).copy(), np.asarray(mine.origin_buffer[i], np.float64))
                 for i in range(B) if mine.origin_buffer[i] is not None]
        gathered = [None] * world
        dist.all_gather_object(gathered, local)
        # ... but a rank keeps only the rows it owns under the SLOT's origin; everything else is garbage here
        mirrors = []
        for slots in gathered:
            for idx, org in slots:
                own = (ys + int(org[1])) % world == rank
                m = rng.integers(-40, 40, size=idx.shape).astype(idx.dtype)
                m[:, own, :] = idx[:, own, :]
                mirrors.append((m.reshape(-1), org))
        occ, passes = fold_slots(mirrors, origin, S, Z)
        total = np.where(occ, OCC, np.minimum(passes, OCC - 1)).astype(np.int32)
        codes = finish_codes(total, prev_codes, prev_origin, origin, S, Z)
        own_now = (ys + int(origin[1])) % world == rank
        shard = rng.integers(-40, 40, size=codes.shape).astype(np.int32)     # rows of other ranks: garbage in MY state
        shard[:, own_now, :] = codes[:, own_now, :]
        prev_codes, prev_origin = shard, origin
        # every rank delivers its rows; the assembly is what the 2-D pushes replicate
        parts = [None] * world
        dist.all_gather_object(parts, (own_now, shard[:, own_now, :]))
        full = np.zeros((Z, S, S), np.int32)
        seen = np.zeros(S, int)
        for o_rows, data in parts:
            full[:, o_rows, :] = data
            seen += o_rows
        ok &= bool((seen == 1).all())                           # every row has exactly one owner
        if rank == 0:
            for r in range(world):
                ref.Process_pointcloud(*fr[step][r])
            ref.combine_maps()
            want = np.where(ref.combined_index_map >= 0, 0, ref.combined_index_map).reshape(Z, S, S)
            ok &= bool(np.array_equal(full, want))
    result[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_mirrored_row_sharded_protocol_gloo():
    world = 2
    with mp.Manager() as m:
        result = m.dict()
        mp.spawn(mirror_worker, args=(world, 29641, result), nprocs=world, join=True)
        assert dict(result) == {0: True, 1: True}


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
