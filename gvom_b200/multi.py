"""Multi-GPU G-VOM: independent sensor streams per GPU, one combined map.

README of the reference (README.md:49) allows several sensors to feed one Gvom;
the per-scan work (Process_pointcloud) of different sensors is independent until
combine_maps() merges the ring buffer.  That is where the path shards
(SURVEY.md 8e): rank g owns sensor stream g and its own ring slots on its own B200.

Default (xy_size % 256 == 0; gvom_mirror.cuh) -- MIRRORED RING SLOTS, ROW-SHARDED COMBINE:
  * the combined map is sharded by world rows: rank r owns the rows y with (y + origin_y) % world == r
  * every Process_pointcloud ends with a push of the new scan to the owners of its rows (posted NVLink stores into
    torch symmetric memory), so every rank holds local copies of all ranks' ring slots for the rows it owns
  * combine_maps(): one flag exchange on the device, the single-GPU merge kernels over the own rows from LOCAL memory,
    then two pushes of 2-D data (heights, finished maps).  No NCCL call, no host synchronisation.

Generic exchange (any grid size; GVOM_MULTI_MIRROR=0) -- PARTIAL MERGE + REPLICATED FINISH:
  1. every rank folds its OWN slots into the common frame (gvom_combine_partial):
     a dense int32 code grid (occupied flag | summed pass count) + compact records
  2. "p2p": the grids / records live in symmetric memory; one device-side barrier, then the finishing kernels read
     the peers' buffers directly over NVLink.  "nccl": all-reduce(sum) of the grids + all-gather of records and a
     header through torch.distributed (the fallback when symmetric memory cannot be set up).
  3. every rank finishes the combine (gvom_combine_finish); the combined state is replicated.

(Built, measured and removed over the two rounds: every rank merging every rank's slots read in place over NVLink
("direct") or from bulk-copied mirrors ("pull"); a plane-sharded finish with full re-assembly; a row-sharded finish
fed by partial merges.  DESIGN.md section 6 has the numbers.)

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

    def __init__(self, *args, group=None, torch_stream=None, exchange="auto", mirror="auto", **kw):
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
        # mirrored ring slots + row-sharded combine (default where it applies: xy_size % 256 == 0, at most 64 ring slots
        # over all ranks): every scan is pushed to the owners of its world rows at Process_pointcloud time, the combine
        # merges the own rows from local mirrors; only 2-D maps are replicated.  GVOM_MULTI_MIRROR=0 selects the generic
        # exchange (partial merge + replicated finish) for A/B runs.
        if mirror == "auto":
            mirror = os.environ.get("GVOM_MULTI_MIRROR", "1") != "0"
        self._mirror = bool(mirror) and self.xy_size % 256 == 0 and self.world * self.buffer_size <= 64
        self._rows = self._mirror                 # (the 3-D combined state is sharded by world rows)
        if exchange in ("auto", "p2p"):
            try:
                self._init_p2p()
                self.exchange = "p2p"
            except Exception as ex:          # symmetric memory unavailable: fall back (all ranks agree below)
                self._p2p_error = repr(ex)
                self._mirror = self._rows = False
                if exchange == "p2p":
                    raise
        else:
            self._mirror = self._rows = False
        # every rank must use the same exchange
        ok = torch.tensor([1 if self.exchange == "p2p" else 0], device=self._dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            if self._mirror:
                raise RuntimeError("MultiGpuGvom: symmetric memory came up on some ranks only, after the mirrors were attached")
            self.exchange = "nccl"
            self._init_nccl()

    def _grow(self, n_points):
        raise RuntimeError(f"MultiGpuGvom: a scan of {n_points} points exceeds max_points={self.max_points}; the exchange "
                           "buffers are sized at construction (pass max_points=...)")

    # ------------------------------------------------------------------ buffers
    def _layout(self):
        """byte offsets inside one exchange set.  Generic exchange: grid | group mask | records | count(+pad) | header |
        flags.  Mirrored combine: "heights" flags | "results" flags | 2-D block."""
        if self._mirror:
            nb = C.c_uint64(0)
            check(self._L.gvom_rows_block_size(self._h, C.byref(nb)), "gvom_rows_block_size")
            self._o_fh, self._o_fr, self._o_b2d = 0, 512, 1024
            return 0, 0, 0, 0, 0, self._o_b2d + int(nb.value) + 256, 0
        V, cap = self.voxel_count, self._rec_cap
        o_grid = 0
        o_msk = (o_grid + 4 * V + 255) & ~255
        o_rec = (o_msk + 4 * (V // 256 + 2) + 255) & ~255
        o_cnt = (o_rec + 4 * RECORD_FLOATS * cap + 255) & ~255
        o_hdr = o_cnt + 256
        o_flg = o_hdr + 8 * HEADER_DOUBLES + 256          # flags[rank] int32: rank's epoch, written by that rank
        total = o_flg + 4 * 64 + 256
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
            if self._mirror:
                K = GvomRowsLinks()
                K.rank, K.nranks = self.rank, self.world
                for r, p in enumerate(ptrs):
                    K.blocks2d[r] = p + self._o_b2d
                    K.heights_slots[r] = p + self._o_fh + 4 * self.rank
                    K.results_slots[r] = p + self._o_fr + 4 * self.rank
                K.heights_flags, K.results_flags = ptrs[self.rank] + self._o_fh, ptrs[self.rank] + self._o_fr
                self._sets.append({"t": t, "hdl": hdl, "me": ptrs[self.rank], "links": K})
                continue
            o_grid, o_msk, o_rec, o_cnt, o_hdr, _, o_flg = self._off
            self._sets.append({
                "t": t, "hdl": hdl, "me": ptrs[self.rank],
                # per-rank pointer tables handed to gvom_combine_finish (constant: built once)
                "grids": _ptr_array([p + o_grid for p in ptrs]), "masks": _ptr_array([p + o_msk for p in ptrs]),
                "recs": _ptr_array([p + o_rec for p in ptrs]), "cnts": _ptr_array([p + o_cnt for p in ptrs]),
                "hdr_view": t[o_hdr:o_hdr + 8 * HEADER_DOUBLES].view(torch.float64),
                # my flag slot in every rank's block (signal) / all ranks' slots in my block (wait)
                "signal": _ptr_array([p + o_flg + 4 * self.rank for p in ptrs]), "wait": ptrs[self.rank] + o_flg,
            })
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
        epoch = self._calls
        have = L.gvom_newest_origin(self._h, self._org_in) != GVOM_NO_DATA
        if self._mirror:
            return self._combine_mirror(device_outputs, wait, X, epoch, have)
        t, hdl, me = X["t"], X["hdl"], X["me"]
        o_grid, o_msk, o_rec, o_cnt, o_hdr, _, o_flg = self._off
        with torch.cuda.stream(self._tstream):
            hh = self._hdr_host                      # header: only read by ranks that have no scan yet (start-up)
            hh[0] = 1.0 if have else 0.0
            hh[2], hh[3], hh[4] = (self._org_in[0], self._org_in[1], self._org_in[2]) if have else (0.0, 0.0, 0.0)
            X["hdr_view"].copy_(hh, non_blocking=True)
            if have:
                # partial kernels, then the signal kernel: "rank `me`, combine `epoch`: done" into every rank's block
                check(L.gvom_combine_partial(self._h, self._org_in, me + o_grid, me + o_msk, me + o_rec, self._rec_cap,
                                             me + o_cnt, X["signal"], self.world, epoch, self._stream), "gvom_combine_partial")
            else:
                # start-up only: this rank has no scan yet.  Publish an empty grid, signal, then wait (on the host)
                # for the others and adopt the origin of a rank that has data.
                t[o_grid:o_rec].zero_()              # empty grid, empty group mask
                t[o_cnt:o_cnt + 4].zero_()
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
            outs, optr, mem = self._outputs(device_outputs)
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
