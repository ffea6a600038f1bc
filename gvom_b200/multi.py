"""Multi-GPU G-VOM: independent sensor streams per GPU, one combined map.

README of the reference (README.md:49) allows several sensors to feed one Gvom;
the per-scan work (Process_pointcloud) of different sensors is independent until
combine_maps() merges the ring buffer.  That is where the path shards
(SURVEY.md 8e): rank g owns sensor stream g and its own ring slots on its own
B200, and the exchange happens only at combine time:

  1. every rank folds its OWN slots into the common frame (gvom_combine_partial):
     a dense int32 code grid (occupied flag | summed pass count) + compact records
  2. exchange, one of
       "p2p"  (default): the grids / records live in torch symmetric memory, i.e.
              every rank's buffers are mapped into every other rank over NVLink.
              One device-side barrier, then the finishing kernels read the peers'
              buffers directly -- compute and "collective" are the same kernels,
              no NCCL launch, no host synchronisation, no staging copy.
       "nccl": all-reduce(sum) of the grids + all-gather of records and a header
              through torch.distributed (the baseline; also the fallback when
              symmetric memory cannot be set up).
  3. every rank finishes the combine (gvom_combine_finish) -- redundantly, which is
     cheaper than broadcasting the result and keeps the "previous combined map"
     state replicated, so any rank can serve the maps.

(Two replicated alternatives were built and measured in round 1 -- every rank merging every rank's slots read in
place over NVLink ("direct"), or from bulk-copied mirrors ("pull") -- lost to partial + finish already at N = 2
 and were removed.)

The result equals a single Gvom holding all ranks' slots: occupancy (OR), pass
sums, hit/total sums and min heights are order independent in the reference's
merge rules; moments agree to float32 rounding (tests/test_multi_gpu.py).
All sensors must share the vehicle ego position (as in gvom_ros.py, where the
origin comes from odometry, not from the sensor pose).
"""
import ctypes as C
import time

import numpy as np

from ._lib import GVOM_DEVICE, GVOM_HOST, GVOM_NO_DATA, GVOM_NONE, RECORD_FLOATS, GvomRowsLinks, check
from .gvom import Gvom

HEADER_DOUBLES = 8          # [valid, count, ox, oy, oz, pad...]


def merge_headers(headers):
    """headers: (world, >=5) float64 rows [valid, count, ox, oy, oz] -> (origin or None, counts int32).
    Pure host logic (covered by the CPU gloo test)."""
    headers = np.asarray(headers, dtype=np.float64)
    headers = headers.reshape(-1, headers.shape[-1])[:, :5]
    valid = headers[:, 0] > 0
    counts = np.where(valid, headers[:, 1], 0).astype(np.int32)
    if not valid.any():
        return None, counts
    org = headers[valid][0, 2:5]
    if not np.all(headers[valid][:, 2:5] == org):
        raise RuntimeError("multi-GPU combine: ranks disagree on the map origin (sensors must share the ego position)")
    return org.copy(), counts


MIRROR_ENTRY_INTS = 16      # include/gvom_b200.h: slot-table entry of the mirrored exchange


def newest_origin(table):
    """table: (ranks, slots, 16) int32 slot-table entries {seq (0: empty), ox, oy, oz, cells, ..., ego xyz float64 at
    ints 8..13} of the mirrored exchange -> (origin, ego) of the newest scan of the first rank that holds one, or
    (None, None).  Pure host logic (start-up path only)."""
    table = np.ascontiguousarray(table, dtype=np.int32)
    for g in range(table.shape[0]):
        seq = table[g, :, 0]
        if (seq > 0).any():
            e = table[g, int(np.argmax(seq))]
            return e[1:4].astype(np.float64), e[8:14].copy().view(np.float64).copy()
    return None, None


def _ptr_array(ptrs):
    return (C.c_void_p * len(ptrs))(*[C.c_void_p(int(p)) for p in ptrs])


class MultiGpuGvom(Gvom):
    """One rank of a multi-GPU Gvom.  Same API as Gvom; combine_maps() is collective
    (every rank must call it) and returns the same maps on every rank."""

    def __init__(self, *args, group=None, torch_stream=None, exchange="auto", sharded="auto", rows="auto", mirror="auto", **kw):
        import os
        import torch
        import torch.distributed as dist
        self._dist, self._group = dist, group
        exchange = os.environ.get("GVOM_MULTI", exchange)
        if torch_stream is None:
            dev = kw.get("device")
            torch_stream = torch.cuda.Stream(device=torch.cuda.current_device() if dev is None else dev)
        # the library kernels and the torch ops of the exchange (copies, fills, NCCL) must share ONE stream
        if kw.get("stream") is not None and int(kw["stream"]) != int(torch_stream.cuda_stream):
            raise ValueError("MultiGpuGvom: `stream` and `torch_stream` name different CUDA streams")
        kw["stream"] = torch_stream.cuda_stream
        self._tstream = torch_stream
        super().__init__(*args, **kw)
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._dev = torch.device(f"cuda:{self.device}")
        # a rank's records: at most one per occupied voxel of its slots
        self._rec_cap = int(min(self.voxel_count, self.buffer_size * self.max_points))
        self._org_in = (C.c_double * 3)()
        self._calls = 0
        self.exchange = None
        # sharded finish (each rank merges 1/world of the z-planes) needs whole warps per plane
        # measured on 8x B200 (profiles/): the sharded finish wins from ~6 ranks up; below that the extra assembly
        # pass costs more than the finishing work it saves, so the replicated finish is used
        if sharded == "auto":
            sharded = self.world >= 6
        self._sharded = bool(sharded) and self.xy_size % 16 == 0
        # row-sharded finish (default where it applies: xy_size % 256 == 0): every rank merges only the world rows it owns
        # and keeps the 3-D state of those rows; only 2-D maps are replicated.  GVOM_MULTI_ROWS=0 selects the older
        # finishes (replicated / plane-sharded with full assembly) for A/B runs.
        if rows == "auto":
            rows = os.environ.get("GVOM_MULTI_ROWS", "1") != "0"
        self._rows = bool(rows) and self.xy_size % 256 == 0
        if self._rows:
            self._sharded = False
        # mirrored ring slots (default with the row-sharded finish): every scan is pushed to the owners of its rows at
        # Process_pointcloud time, the combine merges from local mirrors -- no partial merge, no encoded grids.
        # GVOM_MULTI_MIRROR=0 selects the partial-merge exchange for A/B runs.
        if mirror == "auto":
            mirror = os.environ.get("GVOM_MULTI_MIRROR", "1") != "0"
        self._mirror = bool(mirror) and self._rows and self.world * self.buffer_size <= 64
        nb = C.c_uint64(0)
        check(self._L.gvom_rows_block_size(self._h, C.byref(nb)), "gvom_rows_block_size")
        self._b2d_bytes = int(nb.value)
        ccap = min(self.voxel_count, 4 * self.max_points * (self.buffer_size + 1))
        self._res_cap = int(min(self.voxel_count, max(1 << 18, 4 * ccap // self.world)))
        if exchange in ("auto", "p2p"):
            try:
                self._init_p2p()
                self.exchange = "p2p"
            except Exception as ex:          # symmetric memory unavailable: fall back (all ranks agree below)
                self._p2p_error = repr(ex)
                if exchange == "p2p":
                    raise
        # every rank must use the same exchange
        ok = torch.tensor([1 if self.exchange == "p2p" else 0], device=self._dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            self.exchange = "nccl"
            self._init_nccl()

    def _grow(self, n_points):
        raise RuntimeError(f"MultiGpuGvom: a scan of {n_points} points exceeds max_points={self.max_points}; the exchange "
                           "buffers are sized at construction (pass max_points=...)")

    # ------------------------------------------------------------------ buffers
    def _layout(self):
        """byte offsets inside one exchange set: grid | group mask | records | count(+pad) | header"""
        V, cap = self.voxel_count, self._rec_cap
        o_grid = 0
        o_msk = (o_grid + 4 * V + 255) & ~255
        o_rec = (o_msk + 4 * (V // 256 + 2) + 255) & ~255
        o_cnt = (o_rec + 4 * RECORD_FLOATS * cap + 255) & ~255
        o_hdr = o_cnt + 256
        o_flg = o_hdr + 8 * HEADER_DOUBLES + 256          # flags[rank] int32: rank's epoch, written by that rank
        total = o_flg + 4 * 64 + 256
        # sharded finish: slab-done flags, result count, result index map, result cells (68 B per row)
        self._o_flg2 = total
        self._o_rcnt = self._o_flg2 + 4 * 64 + 256
        self._o_rmap = self._o_rcnt + 4 * 64 + 256
        self._o_rcel = (self._o_rmap + 4 * V + 255) & ~255
        if self._sharded:
            total = self._o_rcel + 68 * self._res_cap + 256
        if self._rows:
            # row-sharded finish: {epoch, origin} headers of the partial results, "heights" / "results" flags, the 2-D block
            self._o_hdr4 = (total + 255) & ~255
            self._o_fh = self._o_hdr4 + 16 * 64 + 256
            self._o_fr = self._o_fh + 4 * 64 + 256
            self._o_b2d = (self._o_fr + 4 * 64 + 255 + 256) & ~255
            total = self._o_b2d + self._b2d_bytes + 256
        return o_grid, o_msk, o_rec, o_cnt, o_hdr, total, o_flg

    def _init_p2p(self):
        import torch.distributed._symmetric_memory as symm_mem
        torch, dist = self._torch, self._dist
        group = self._group if self._group is not None else dist.group.WORLD
        self._off = self._layout()
        total = self._off[5]
        self._sets = []
        for _ in range(2):                   # double buffered: one device barrier per combine is enough
            t = symm_mem.empty(total, dtype=torch.uint8, device=self._dev)
            hdl = symm_mem.rendezvous(t, group)
            t.zero_()
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            o_grid, o_msk, o_rec, o_cnt, o_hdr, _, o_flg = self._off
            self._sets.append({
                "t": t, "hdl": hdl, "me": ptrs[self.rank],
                # per-rank pointer tables handed to gvom_combine_finish (constant: built once)
                "grids": _ptr_array([p + o_grid for p in ptrs]), "masks": _ptr_array([p + o_msk for p in ptrs]),
                "recs": _ptr_array([p + o_rec for p in ptrs]), "cnts": _ptr_array([p + o_cnt for p in ptrs]),
                "hdr_view": t[o_hdr:o_hdr + 8 * HEADER_DOUBLES].view(torch.float64),
                # my flag slot in every rank's block (signal) / all ranks' slots in my block (wait)
                "signal": _ptr_array([p + o_flg + 4 * self.rank for p in ptrs]), "wait": ptrs[self.rank] + o_flg,
                "signal2": _ptr_array([p + self._o_flg2 + 4 * self.rank for p in ptrs]), "wait2": ptrs[self.rank] + self._o_flg2,
                "rmaps": _ptr_array([p + self._o_rmap for p in ptrs]), "rcells": _ptr_array([p + self._o_rcel for p in ptrs]),
                # count table (64 ints) in every rank's block: entry r is pushed by rank r
                "cslots": _ptr_array([p + self._o_rcnt + 4 * self.rank for p in ptrs]), "ctable": ptrs[self.rank] + self._o_rcnt,
            })
            if self._rows:
                K = GvomRowsLinks()
                K.rank, K.nranks, K.record_capacity = self.rank, self.world, self._rec_cap
                for r, p in enumerate(ptrs):
                    K.code_grids[r], K.group_masks[r], K.records[r] = p + o_grid, p + o_msk, p + o_rec
                    K.blocks2d[r] = p + self._o_b2d
                    K.heights_slots[r] = p + self._o_fh + 4 * self.rank
                    K.results_slots[r] = p + self._o_fr + 4 * self.rank
                me_p = ptrs[self.rank]
                K.partial_headers, K.heights_flags, K.results_flags = me_p + self._o_hdr4, me_p + self._o_fh, me_p + self._o_fr
                self._sets[-1]["links"] = K
                self._sets[-1]["hdr4"] = _ptr_array([p + self._o_hdr4 + 16 * self.rank for p in ptrs])
        if self._mirror:
            nb = C.c_uint64(0)
            check(self._L.gvom_mirror_block_size(self._h, self.world, C.byref(nb)), "gvom_mirror_block_size")
            self._mir_t = symm_mem.empty(int(nb.value), dtype=torch.uint8, device=self._dev)
            self._mir_hdl = symm_mem.rendezvous(self._mir_t, group)
            check(self._L.gvom_mirror_attach(self._h, self.rank, self.world, _ptr_array([int(p) for p in self._mir_hdl.buffer_ptrs])),
                  "gvom_mirror_attach")
        torch.cuda.synchronize(self._dev)
        dist.barrier(group=self._group)          # nobody scans (pushes) before every rank's mirrors are initialised
        self._hdr_host = torch.zeros(HEADER_DOUBLES, dtype=torch.float64).pin_memory()

    def _init_nccl(self):
        torch = self._torch
        self._grid = torch.empty(self.voxel_count, dtype=torch.int32, device=self._dev)
        self._records = torch.empty((self._rec_cap, RECORD_FLOATS), dtype=torch.float32, device=self._dev)
        self._count = torch.zeros(1, dtype=torch.int32, device=self._dev)
        self._header = torch.zeros(HEADER_DOUBLES, dtype=torch.float64, device=self._dev)
        self._headers = torch.zeros((self.world, HEADER_DOUBLES), dtype=torch.float64, device=self._dev)
        self._counts_dev = torch.zeros(self.world, dtype=torch.int32, device=self._dev)

    # ------------------------------------------------------------------ outputs
    def _outputs(self, device_outputs):
        torch, S = self._torch, self.xy_size
        if device_outputs:
            ti = torch.empty((3, S, S), dtype=torch.int32, device=self._dev)
            rough = torch.empty((S, S), dtype=torch.float64, device=self._dev)
            return (ti[0], ti[1], rough, ti[2]), (ti[0].data_ptr(), ti[1].data_ptr(), rough.data_ptr(), ti[2].data_ptr()), GVOM_DEVICE
        pos, neg, rough, vis = self._out_arrays()
        return (pos, neg, rough, vis), (pos.ctypes.data, neg.ctypes.data, rough.ctypes.data, vis.ctypes.data), GVOM_HOST

    def combine_maps(self, device_outputs=False, wait=True):
        """Collective.  wait=False (device outputs, row-sharded exchange only): return once everything is enqueued; the
        tensors are valid in stream order on the Gvom's stream."""
        self._calls += 1
        if self.exchange == "p2p":
            return self._combine_p2p(device_outputs, wait)
        return self._combine_nccl(device_outputs)

    # ------------------------------------------------------------------ peer-to-peer exchange
    def _combine_p2p(self, device_outputs, wait=True):
        torch, L = self._torch, self._L
        X = self._sets[self._calls & 1]
        t, hdl, me = X["t"], X["hdl"], X["me"]
        o_grid, o_msk, o_rec, o_cnt, o_hdr, _, o_flg = self._off
        epoch = self._calls
        have = L.gvom_newest_origin(self._h, self._org_in) != GVOM_NO_DATA
        if self._mirror:
            return self._combine_mirror(device_outputs, wait, X, epoch, have)
        with torch.cuda.stream(self._tstream):
            hh = self._hdr_host                      # header: only read by ranks that have no scan yet (start-up)
            hh[0] = 1.0 if have else 0.0
            hh[2], hh[3], hh[4] = (self._org_in[0], self._org_in[1], self._org_in[2]) if have else (0.0, 0.0, 0.0)
            X["hdr_view"].copy_(hh, non_blocking=True)
            if have and self._rows:
                # partial kernels, then the header kernel: {epoch, origin} into every rank's block
                check(L.gvom_combine_partial_header(self._h, self._org_in, me + o_grid, me + o_msk, me + o_rec, self._rec_cap,
                                                    me + o_cnt, X["hdr4"], self.world, epoch, self._stream),
                      "gvom_combine_partial_header")
            elif have:
                # partial kernels, then the signal kernel: "rank `me`, combine `epoch`: done" into every rank's block
                check(L.gvom_combine_partial(self._h, self._org_in, me + o_grid, me + o_msk, me + o_rec, self._rec_cap,
                                             me + o_cnt, X["signal"], self.world, epoch, self._stream), "gvom_combine_partial")
            else:
                # start-up only: this rank has no scan yet.  Publish an empty grid, signal, then wait (on the host)
                # for the others and adopt the origin of a rank that has data.
                t[o_grid:o_rec].zero_()              # empty grid, empty group mask
                t[o_cnt:o_cnt + 4].zero_()
                if self._rows:       # header {epoch, no origin}
                    mine = torch.tensor([epoch, 0x7fffffff, 0, 0], dtype=torch.int32, device=self._dev)
                    for r in range(self.world):
                        hdl.get_buffer(r, (64 * 4,), torch.int32, self._o_hdr4 // 4)[4 * self.rank:4 * self.rank + 4].copy_(mine)
                    flags = t[self._o_hdr4:self._o_hdr4 + 16 * 64].view(torch.int32)[:4 * self.world:4]
                else:
                    for r in range(self.world):
                        hdl.get_buffer(r, (64,), torch.int32, o_flg // 4)[self.rank:self.rank + 1].fill_(epoch)
                    flags = t[o_flg:o_flg + 4 * 64].view(torch.int32)[:self.world]
                self._tstream.synchronize()
                while int(flags.min().item()) < epoch:
                    time.sleep(1e-4)
                heads = np.stack([hdl.get_buffer(r, (HEADER_DOUBLES,), torch.float64, o_hdr // 8).cpu().numpy()
                                  for r in range(self.world)])
                origin, _ = merge_headers(heads)
                if origin is None:
                    print("ERROR: No data in buffer")
                    return None
                for k in range(3):
                    self._org_in[k] = float(origin[k])
            if self._rows:
                # own world rows: merge + cells + columns; heights and finished maps are pushed to every rank
                outs, optr, mem = self._outputs(device_outputs)
                phases = 7 if (wait or not device_outputs) else 15
                check(L.gvom_combine_finish_rows(self._h, self._org_in, C.byref(X["links"]), epoch, phases, self._org_c,
                                                 optr[0], optr[1], optr[2], optr[3], mem, self._stream),
                      "gvom_combine_finish_rows")
                if phases == 15:
                    self._hold_until_consumed(outs)      # the stream may still be writing them when the caller drops them
                pos, neg, rough, vis = outs
                return (np.array([self._org_c[0], self._org_c[1], self._org_c[2]]), pos, neg, rough, vis)
            outs, optr, mem = self._outputs(device_outputs)
            if self._sharded:
                # every rank finishes 1/world of the planes, publishes them, and assembles the full map from all ranks
                check(L.gvom_combine_finish_sharded(self._h, self._org_in, self.rank, self.world, X["grids"], X["masks"],
                                                    X["recs"], self._rec_cap, X["wait"], X["rmaps"], X["rcells"], X["cslots"],
                                                    X["ctable"], self._res_cap, X["signal2"], X["wait2"], epoch, 3, self._org_c,
                                                    optr[0], optr[1], optr[2], optr[3], mem, self._stream),
                      "gvom_combine_finish_sharded")
                pos, neg, rough, vis = outs
                return (np.array([self._org_c[0], self._org_c[1], self._org_c[2]]), pos, neg, rough, vis)
            # the finishing merge kernel waits on the flag slots itself, then reads the peers' buffers over NVLink
            check(L.gvom_combine_finish(self._h, self._org_in, X["grids"], X["masks"], self.world, X["recs"], X["cnts"],
                                        self.world, self._rec_cap, X["wait"], epoch, self._org_c, optr[0], optr[1],
                                        optr[2], optr[3], mem, self._stream), "gvom_combine_finish")
        pos, neg, rough, vis = outs
        return (np.array([self._org_c[0], self._org_c[1], self._org_c[2]]), pos, neg, rough, vis)

    # ------------------------------------------------------------------ mirrored ring slots
    def _combine_mirror(self, device_outputs, wait, X, epoch, have):
        """Row-sharded combine over the local mirrors of every rank's ring slots (gvom_mirror_attach): the scans were
        pushed when they were made; here one warp publishes / awaits the epoch flags and builds the source list."""
        torch, L = self._torch, self._L
        extra = 0
        with torch.cuda.stream(self._tstream):
            if not have:
                # start-up only: this rank has no scan yet.  Publish the flag, wait (on the host) for the others and
                # adopt the origin of a rank that has data (the slot table the pushes carried).
                check(L.gvom_combine_finish_rows(self._h, self._org_in, C.byref(X["links"]), epoch, 16, self._org_c,
                                                 None, None, None, None, GVOM_NONE, self._stream), "gvom_combine_finish_rows")
                self._tstream.synchronize()
                flags = self._mir_t[:4 * self.world].view(torch.int32)
                while int(flags.min().item()) < epoch:
                    time.sleep(1e-4)
                nt = self.world * self.buffer_size * MIRROR_ENTRY_INTS
                table = self._mir_t[256:256 + 4 * nt].view(torch.int32).cpu().numpy().reshape(self.world, self.buffer_size, MIRROR_ENTRY_INTS)
                origin, ego = newest_origin(table)
                if origin is None:
                    print("ERROR: No data in buffer")
                    return None
                for k in range(3):
                    self._org_in[k] = float(origin[k])
                check(L.gvom_adopt_ego(self._h, (C.c_double * 3)(*[float(v) for v in ego])), "gvom_adopt_ego")
                extra = 32
            outs, optr, mem = self._outputs(device_outputs)
            phases = (7 if (wait or not device_outputs) else 15) | extra
            check(L.gvom_combine_finish_rows(self._h, self._org_in, C.byref(X["links"]), epoch, phases, self._org_c,
                                             optr[0], optr[1], optr[2], optr[3], mem, self._stream),
                  "gvom_combine_finish_rows")
            if phases & 8:
                self._hold_until_consumed(outs)
        pos, neg, rough, vis = outs
        return (np.array([self._org_c[0], self._org_c[1], self._org_c[2]]), pos, neg, rough, vis)

    # ------------------------------------------------------------------ NCCL exchange
    def _combine_nccl(self, device_outputs):
        torch, dist, L = self._torch, self._dist, self._L
        have = L.gvom_newest_origin(self._h, self._org_in) != GVOM_NO_DATA
        org = [self._org_in[0], self._org_in[1], self._org_in[2]] if have else [0.0, 0.0, 0.0]
        with torch.cuda.stream(self._tstream):
            if have:
                check(L.gvom_combine_partial(self._h, self._org_in, self._grid.data_ptr(), None, self._records.data_ptr(),
                                             self._rec_cap, self._count.data_ptr(), None, 0, 0, self._stream),
                      "gvom_combine_partial")
            else:
                self._grid.zero_()
                self._count.zero_()
            hdr = [1.0 if have else 0.0, 0.0] + org + [0.0] * (HEADER_DOUBLES - 5)
            self._header.copy_(torch.tensor(hdr, dtype=torch.float64), non_blocking=False)
            self._header[1:2] = self._count.to(torch.float64)
            dist.all_gather_into_tensor(self._headers, self._header, group=self._group)
            dist.all_reduce(self._grid, op=dist.ReduceOp.SUM, group=self._group)
            headers = self._headers.cpu().numpy()            # one host sync: sizes of the record exchange
            origin, counts = merge_headers(headers)
            if origin is None:
                print("ERROR: No data in buffer")
                return None
            if (counts > self._rec_cap).any():
                raise RuntimeError("multi-GPU combine: a rank has more occupied voxels than its record capacity")
            maxc = max(1, int(counts.max()))
            gathered = torch.empty((self.world, maxc, RECORD_FLOATS), dtype=torch.float32, device=self._dev)
            dist.all_gather_into_tensor(gathered, self._records[:maxc], group=self._group)
            self._counts_dev.copy_(torch.from_numpy(counts))
            for k in range(3):
                self._org_in[k] = float(origin[k])
            outs, optr, mem = self._outputs(device_outputs)
            grids = _ptr_array([self._grid.data_ptr()])
            recs = _ptr_array([gathered.data_ptr() + 4 * RECORD_FLOATS * maxc * r for r in range(self.world)])
            cnts = _ptr_array([self._counts_dev.data_ptr() + 4 * r for r in range(self.world)])
            check(L.gvom_combine_finish(self._h, self._org_in, grids, None, 1, recs, cnts, self.world, maxc, None, 0, self._org_c,
                                        optr[0], optr[1], optr[2], optr[3], mem, self._stream),
                  "gvom_combine_finish")
        pos, neg, rough, vis = outs
        return (np.array([self._org_c[0], self._org_c[1], self._org_c[2]]), pos, neg, rough, vis)
