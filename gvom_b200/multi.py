"""Multi-GPU G-VOM: independent sensor streams per GPU, one combined map.

README of the reference (README.md:49) allows several sensors to feed one Gvom;
the per-scan work (Process_pointcloud) of different sensors is independent until
combine_maps() merges the ring buffer.  That is where the path shards
(SURVEY.md 8e): rank g owns sensor stream g and its own ring slots on its own
B200, and the exchange happens only at combine time:

  1. every rank folds its OWN slots into the common frame (gvom_combine_partial):
     a dense int32 code grid (occupied flag | summed pass count) + compact records
  2. NCCL over NVLink (torch.distributed): all-reduce(sum) of the grids,
     all-gather of the records and of a small header (count, origin)
  3. every rank finishes the combine redundantly (gvom_combine_finish) -- cheaper
     than broadcasting the result and it keeps the "previous combined map" state
     replicated, so any rank can serve the maps.

The result equals a single Gvom holding all ranks' slots: occupancy (OR), pass
sums, hit/total sums and min heights are order independent in the reference's
merge rules; moments agree to float32 rounding (tests/test_multi_gpu.py).
All sensors must share the vehicle ego position (as in gvom_ros.py, where the
origin comes from odometry, not from the sensor pose).
"""
import ctypes as C

import numpy as np

from ._lib import GVOM_DEVICE, GVOM_HOST, GVOM_NO_DATA, RECORD_FLOATS, check
from .gvom import Gvom


def merge_headers(headers):
    """headers: (world, 5) float64 rows [valid, count, ox, oy, oz] -> (origin or None, counts int32).
    Pure host logic (covered by the CPU gloo test)."""
    headers = np.asarray(headers, dtype=np.float64).reshape(-1, 5)
    valid = headers[:, 0] > 0
    counts = np.where(valid, headers[:, 1], 0).astype(np.int32)
    if not valid.any():
        return None, counts
    org = headers[valid][0, 2:5]
    if not np.all(headers[valid][:, 2:5] == org):
        raise RuntimeError("multi-GPU combine: ranks disagree on the map origin (sensors must share the ego position)")
    return org.copy(), counts


class MultiGpuGvom(Gvom):
    """One rank of a multi-GPU Gvom.  Same API as Gvom; combine_maps() is collective
    (every rank must call it) and returns the same maps on every rank."""

    def __init__(self, *args, group=None, torch_stream=None, **kw):
        import torch
        import torch.distributed as dist
        self._dist, self._group = dist, group
        if torch_stream is None:
            dev = kw.get("device")
            torch_stream = torch.cuda.Stream(device=torch.cuda.current_device() if dev is None else dev)
            kw["stream"] = torch_stream.cuda_stream
        self._tstream = torch_stream
        super().__init__(*args, **kw)
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        dev = f"cuda:{self.device}"
        # a rank's records: at most one per occupied voxel of its slots
        self._rec_cap = int(min(self.voxel_count, self.buffer_size * self.max_points))
        self._grid = torch.empty(self.voxel_count, dtype=torch.int32, device=dev)
        self._records = torch.empty((self._rec_cap, RECORD_FLOATS), dtype=torch.float32, device=dev)
        self._count = torch.zeros(1, dtype=torch.int32, device=dev)
        self._header = torch.zeros(5, dtype=torch.float64, device=dev)
        self._headers = torch.zeros((self.world, 5), dtype=torch.float64, device=dev)
        self._counts_dev = torch.zeros(self.world, dtype=torch.int32, device=dev)
        self._org_in = (C.c_double * 3)()

    def combine_maps(self, device_outputs=False):
        torch, dist, L = self._torch, self._dist, self._L
        have = L.gvom_newest_origin(self._h, self._org_in) != GVOM_NO_DATA
        org = [self._org_in[0], self._org_in[1], self._org_in[2]] if have else [0.0, 0.0, 0.0]
        with torch.cuda.stream(self._tstream):
            if have:
                check(L.gvom_combine_partial(self._h, self._org_in, self._grid.data_ptr(), self._records.data_ptr(),
                                             self._rec_cap, self._count.data_ptr(), self._stream),
                      "gvom_combine_partial")
            else:
                self._grid.zero_()
                self._count.zero_()
            # header: [valid, count, origin]
            self._header.copy_(torch.tensor([1.0 if have else 0.0, 0.0] + org, dtype=torch.float64), non_blocking=False)
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
            gathered = torch.empty((self.world, maxc, RECORD_FLOATS), dtype=torch.float32, device=self._records.device)
            dist.all_gather_into_tensor(gathered, self._records[:maxc], group=self._group)
            self._counts_dev.copy_(torch.from_numpy(counts))
            for k in range(3):
                self._org_in[k] = float(origin[k])
            S = self.xy_size
            if device_outputs:
                ti = torch.empty((3, S, S), dtype=torch.int32, device=self._records.device)
                rough = torch.empty((S, S), dtype=torch.float64, device=self._records.device)
                ptrs, mem = (ti[0].data_ptr(), ti[1].data_ptr(), rough.data_ptr(), ti[2].data_ptr()), GVOM_DEVICE
                pos, neg, vis = ti[0], ti[1], ti[2]
            else:
                pos, neg, rough, vis = self._out_arrays()
                ptrs, mem = (pos.ctypes.data, neg.ctypes.data, rough.ctypes.data, vis.ctypes.data), GVOM_HOST
            check(L.gvom_combine_finish(self._h, self._org_in, self._grid.data_ptr(), gathered.data_ptr(),
                                        self._counts_dev.data_ptr(), self.world, maxc, self._org_c,
                                        ptrs[0], ptrs[1], ptrs[2], ptrs[3], mem, self._stream),
                  "gvom_combine_finish")
        return (np.array([self._org_c[0], self._org_c[1], self._org_c[2]]), pos, neg, rough, vis)
