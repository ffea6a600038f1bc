"""Drop-in `Gvom` class over the B200 C-ABI (include/gvom_b200.h).

Mirrors the public surface of the reference class (scripts/gvom.py:8-442) that
scripts/gvom_ros.py uses: the 14 positional constructor arguments
(gvom.py:21-22), Process_pointcloud (gvom.py:105), combine_maps (gvom.py:222)
and the three make_debug_* exports (gvom.py:395-442), with the reference's
return types and its "print and return None" error convention.

Python here is only the host shim: argument checking, buffer ownership (torch
tensors own the device and pinned workspaces) and one ctypes call per method.
All computation is in hand-written sm_100a CUDA behind the C-ABI; there is no
Numba, Triton or CPU fallback.
"""
import collections
import ctypes as C
import threading

import numpy as np

from . import _lib
from ._lib import (GRID_NAMES, GVOM_DEVICE, GVOM_F32, GVOM_F64, GVOM_HOST, GVOM_NO_DATA, GvomParams, GvomStats,
                   check)

DEFAULT_MAX_POINTS = 1 << 19          # 524,288: two OS1-128 scans


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("gvom_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


class Gvom:
    """Lidar scans in, 2-D cost maps out (B200-native build of the reference class `Gvom`).

    The 14 positional arguments are the reference's (gvom.py:21-22), same order and meaning: voxel
    resolutions [m] and grid sizes [voxels] in xy and z, number of scans kept in the ring buffer, minimum point
    range, the obstacle / slope thresholds, robot height and radius, lidar height above ground, and the
    neighbourhood radii (voxels) of the per-voxel covariance.
    Keyword-only extras (not in the reference): max_points (initial capacity of one scan; grows on demand),
    device (CUDA ordinal), max_combined_cells, pinned_outputs, stream.
    """

    def __init__(self, xy_resolution, z_resolution, xy_size, z_size, buffer_size, min_distance,
                 positive_obstacle_threshold, negative_obstacle_threshold, slope_obsacle_threshold,
                 robot_height, robot_radius, ground_to_lidar_height, xy_eigen_dist, z_eigen_dist, *,
                 max_points=DEFAULT_MAX_POINTS, device=None, max_combined_cells=0, pinned_outputs=True,
                 stream=None):
        self.xy_resolution, self.z_resolution = xy_resolution, z_resolution
        self.xy_size, self.z_size, self.buffer_size = int(xy_size), int(z_size), int(buffer_size)
        self.min_distance = min_distance
        self.positive_obstacle_threshold = positive_obstacle_threshold
        self.negative_obstacle_threshold = negative_obstacle_threshold
        self.slope_obsacle_threshold = slope_obsacle_threshold
        self.robot_height, self.robot_radius = robot_height, robot_radius
        self.ground_to_lidar_height = ground_to_lidar_height
        self.xy_eigen_dist, self.z_eigen_dist = int(xy_eigen_dist), int(z_eigen_dist)
        self.metrics_count = 10
        self.voxel_count = self.xy_size * self.xy_size * self.z_size
        self.max_points = int(max_points)
        self.pinned_outputs = bool(pinned_outputs)
        self.ego_position = [0, 0, 0]
        self._h = None
        # cudaStream_t (int) all work is enqueued on; None = the handle's own non-blocking stream
        self._stream = None if stream is None else C.c_void_p(int(stream))

        self._L = _lib.lib()                     # raises if the CUDA library is missing
        torch = self._torch = _torch()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._P = GvomParams(float(xy_resolution), float(z_resolution), self.xy_size, self.z_size,
                             self.buffer_size, 0, float(min_distance), float(positive_obstacle_threshold),
                             float(negative_obstacle_threshold), float(slope_obsacle_threshold),
                             float(robot_height), float(robot_radius), float(ground_to_lidar_height),
                             self.xy_eigen_dist, self.z_eigen_dist)
        self._max_combined_cells = int(max_combined_cells)
        self._devstr = f"cuda:{self.device}"
        self._inflight = collections.deque()     # (event, CUDA input tensor) of scans whose kernels may still be running
        self._lock = threading.Lock()            # callback threads (README.md:49: several sensors) share one handle
        self._create()
        self._org_c = (C.c_double * 3)()

    def _create(self):
        """Allocate the workspaces (torch owns the memory; raw pointers cross the ABI) and create the handle."""
        torch = self._torch
        db, hb = C.c_size_t(0), C.c_size_t(0)
        check(self._L.gvom_workspace_size(C.byref(self._P), self.max_points, self._max_combined_cells,
                                          C.byref(db), C.byref(hb)), "gvom_workspace_size")
        self._dev_ws = self._alloc_device_ws(db.value)
        self._host_ws = torch.empty(hb.value, dtype=torch.uint8, pin_memory=True)
        h = C.c_void_p()
        check(self._L.gvom_create(C.byref(self._P), self.max_points, self._max_combined_cells, self.device,
                                  self._dev_ws.data_ptr(), db.value, self._host_ws.data_ptr(), hb.value,
                                  C.byref(h)), "gvom_create")
        self._h = h
        # torch view of the stream the library works on (ordering / lifetime of CUDA-tensor inputs)
        sp = C.c_void_p()
        check(self._L.gvom_get_stream(self._h, C.byref(sp)), "gvom_get_stream")
        raw = self._stream.value if self._stream is not None else sp.value
        self._work_stream = torch.cuda.ExternalStream(raw, device=torch.device(f"cuda:{self.device}"))

    def _grow(self, n_points):
        """A cloud larger than the current capacity arrived (the reference has no limit: it allocates per scan,
        gvom.py:115-131): re-create the handle with twice the room and carry the ring buffer over."""
        blob = self.save_state()
        old_h, old_ws = self._h, (self._dev_ws, self._host_ws)
        self.max_points = max(2 * self.max_points, 1 << (int(n_points) - 1).bit_length())
        self._create()
        self.load_state(blob)
        self._L.gvom_destroy(old_h)
        del old_ws

    def _alloc_device_ws(self, nbytes):
        """Device workspace (torch owns it).  MultiGpuGvom overrides this to place it in symmetric memory."""
        return self._torch.empty(nbytes, dtype=self._torch.uint8, device=f"cuda:{self.device}")

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self):
        if getattr(self, "_h", None):
            self._L.gvom_destroy(self._h)           # synchronises the handle's stream
            self._h = None
            self._inflight.clear()

    # ------------------------------------------------------------------ scans
    def _describe(self, pointcloud):
        """-> (keepalive, pointer, n, stride, dtype code, mem)"""
        torch = self._torch
        if isinstance(pointcloud, torch.Tensor):
            t = pointcloud
            if t.dim() != 2 or t.shape[1] < 3:
                raise ValueError("pointcloud must have shape (N, >=3)")
            if t.dtype not in (torch.float32, torch.float64):
                t = t.to(torch.float64)
            if t.shape[1] > 4:
                t = t[:, :3]
            t = t.contiguous()
            mem = GVOM_DEVICE if t.is_cuda else GVOM_HOST
            if t.is_cuda and t.device.index != self.device:
                t = t.to(f"cuda:{self.device}")
            return t, t.data_ptr(), t.shape[0], t.shape[1], GVOM_F32 if t.dtype == torch.float32 else GVOM_F64, mem
        a = pointcloud if isinstance(pointcloud, np.ndarray) else np.asarray(pointcloud)
        if a.ndim != 2 or a.shape[1] < 3:
            raise ValueError("pointcloud must have shape (N, >=3)")
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        if a.shape[1] > 4:
            a = a[:, :3]
        if not a.flags.c_contiguous:
            a = np.ascontiguousarray(a)
        return a, a.ctypes.data, a.shape[0], a.shape[1], GVOM_F32 if a.dtype == np.float32 else GVOM_F64, GVOM_HOST

    def _order_device_input(self, keep):
        """A CUDA tensor is read in place, asynchronously, on the library's stream: order the tensor's producer
        (torch's current stream) before it and keep the memory from being reused until the kernels are done."""
        torch = self._torch
        cur = torch.cuda.current_stream(keep.device)
        if cur.cuda_stream != self._work_stream.cuda_stream:
            self._work_stream.wait_stream(cur)

    def _hold_until_consumed(self, keep):
        """Keep a reference to a CUDA input tensor (or the temporary made from it) until the kernels that read it
        are done: otherwise torch's caching allocator could hand the memory to another stream's op while the scan
        kernels are still reading it.  (Not tensor.record_stream(): the allocator would later record an event on
        the library's stream, which may no longer exist.)"""
        ev = self._torch.cuda.Event()
        ev.record(self._work_stream)
        q = self._inflight
        q.append((ev, keep))
        while q and q[0][0].query():
            q.popleft()

    def Process_pointcloud(self, pointcloud, ego_position, transform=None):
        """ Imports a pointcloud and processes it into a voxel map then adds the map to the buffer"""
        keep, ptr, n, stride, dt, mem = self._describe(pointcloud)
        e, T, tp = self._ego_and_transform(ego_position, transform)
        if n > self.max_points:
            with self._lock:
                if n > self.max_points:
                    self._grow(n)
        if mem == GVOM_DEVICE:
            self._order_device_input(keep)
        check(self._L.gvom_process_pointcloud(self._h, ptr, n, stride, dt, mem, e, tp, self._stream),
              "gvom_process_pointcloud")
        if mem == GVOM_DEVICE:
            self._hold_until_consumed(keep)
        del keep, T

    process_pointcloud = Process_pointcloud      # spelling used by BASELINE.json

    def _ego_and_transform(self, ego_position, transform):
        self.ego_position = ego_position
        # a fresh buffer per call: concurrent callback threads must not see each other's ego (the C side copies it
        # under the handle's mutex)
        e = (C.c_double * 3)(float(ego_position[0]), float(ego_position[1]), float(ego_position[2]))
        if transform is None:
            return e, None, None
        T = np.ascontiguousarray(transform, dtype=np.float64)
        if T.shape != (4, 4):
            raise ValueError("transform must be 4x4")
        return e, T, T.ctypes.data

    def Process_pointcloud2(self, data, n_points, point_step, ego_position, transform=None, offsets=(0, 4, 8)):
        """Extension: ingest the byte payload of a sensor_msgs/PointCloud2 directly --
        `n_points` records of `point_step` bytes with little-endian float32 x, y, z at byte `offsets`.
        Replaces ros_numpy.point_cloud2.pointcloud2_to_xyz_array(msg) + Process_pointcloud
        (gvom_ros.py:108-109) with the same result: float32 fields widened to float64, NaN / Inf
        records dropped.  `data`: bytes-like, numpy uint8 array, or torch uint8 tensor (CPU or CUDA)."""
        torch = self._torch
        n_points, point_step = int(n_points), int(point_step)
        if isinstance(data, torch.Tensor):
            t = data.contiguous().view(torch.uint8).reshape(-1)
            if t.is_cuda and t.device.index != self.device:
                t = t.to(f"cuda:{self.device}")
            keep, ptr, nbytes, mem = t, t.data_ptr(), t.numel(), GVOM_DEVICE if t.is_cuda else GVOM_HOST
        else:
            a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data).view(np.uint8).reshape(-1)
            keep, ptr, nbytes, mem = a, a.ctypes.data, a.size, GVOM_HOST
        if nbytes < n_points * point_step:
            raise ValueError("PointCloud2 payload shorter than n_points * point_step")
        e, T, tp = self._ego_and_transform(ego_position, transform)
        if n_points > self.max_points:
            with self._lock:
                if n_points > self.max_points:
                    self._grow(n_points)
        if mem == GVOM_DEVICE:
            self._order_device_input(keep)
        check(self._L.gvom_process_pointcloud2(self._h, ptr, n_points, point_step, int(offsets[0]), int(offsets[1]),
                                               int(offsets[2]), mem, e, tp, self._stream), "gvom_process_pointcloud2")
        if mem == GVOM_DEVICE:
            self._hold_until_consumed(keep)
        del keep, T

    def process_pointcloud2_msg(self, msg, ego_position, transform=None):
        """Process_pointcloud2 for an object with the sensor_msgs/PointCloud2 attributes
        (fields[name, offset, datatype], point_step, width, height, data, is_bigendian)."""
        if getattr(msg, "is_bigendian", False):
            raise ValueError("big-endian PointCloud2 payloads are not supported")
        off = {}
        for f in msg.fields:
            if f.name in ("x", "y", "z"):
                if int(f.datatype) != 7:                 # sensor_msgs/PointField.FLOAT32
                    raise ValueError("x / y / z must be FLOAT32 fields")
                off[f.name] = int(f.offset)
        if len(off) != 3:
            raise ValueError("PointCloud2 has no x / y / z fields")
        n = int(msg.width) * int(msg.height)
        if int(msg.height) > 1 and int(msg.row_step) != int(msg.width) * int(msg.point_step):
            raise ValueError("PointCloud2 rows are padded (row_step != width * point_step)")
        return self.Process_pointcloud2(msg.data, n, int(msg.point_step), ego_position, transform,
                                        (off["x"], off["y"], off["z"]))

    # ---------------------------------------------------------------- combine
    def _out_arrays(self):
        """Fresh output arrays (positive, negative, roughness, visibility).  With pinned_outputs
        they are views of ONE pinned block laid out like the device-side result block, so the
        library moves all four maps with a single DMA straight into the returned arrays."""
        S = self.xy_size
        if self.pinned_outputs:
            ro = (12 * S * S + 7) & ~7               # roughness starts 8-byte aligned (odd xy_size)
            blk = self._torch.empty(ro + 8 * S * S, dtype=self._torch.uint8, pin_memory=True).numpy()
            i32 = blk[:12 * S * S].view(np.int32).reshape(3, S, S)
            rough = blk[ro:].view(np.float64).reshape(S, S)
            return i32[0], i32[1], rough, i32[2]
        return (np.empty((S, S), np.int32), np.empty((S, S), np.int32), np.empty((S, S), np.float64),
                np.empty((S, S), np.int32))

    def combine_maps(self, device_outputs=False, wait=True):
        """ Combines all maps in the buffer and processes into 2D maps.
        device_outputs=True (extension) returns torch CUDA tensors instead of numpy arrays; with wait=False the call
        returns as soon as the kernels are enqueued and the tensors are valid in STREAM ORDER on the Gvom's stream
        (like the result of any asynchronous torch op; `synchronize()` or a stream wait before reading them elsewhere)."""
        if device_outputs:
            torch, S = self._torch, self.xy_size
            ro = (12 * S * S + 7) & ~7               # roughness starts 8-byte aligned (odd xy_size)
            blk = torch.empty(ro + 8 * S * S, dtype=torch.uint8, device=self._devstr)   # one allocation, laid out like the result block
            ti = blk[:12 * S * S].view(torch.int32).view(3, S, S)
            rough = blk[ro:].view(torch.float64).view(S, S)
            p0 = blk.data_ptr()
            fn = self._L.gvom_combine_maps if wait else self._L.gvom_combine_maps_async
            rc = check(fn(self._h, self._org_c, p0, p0 + 4 * S * S, p0 + ro, p0 + 8 * S * S, GVOM_DEVICE, self._stream),
                       "gvom_combine_maps")
            if not wait:
                self._hold_until_consumed(blk)       # (the stream may still be writing it when the caller drops it)
            pos, neg, vis = ti[0], ti[1], ti[2]
        else:
            pos, neg, rough, vis = self._out_arrays()
            rc = check(self._L.gvom_combine_maps(self._h, self._org_c, pos.ctypes.data, neg.ctypes.data,
                                                 rough.ctypes.data, vis.ctypes.data, GVOM_HOST, self._stream),
                       "gvom_combine_maps")
        if rc == GVOM_NO_DATA:
            print("ERROR: No data in buffer")
            return None
        origin = np.array([self._org_c[0], self._org_c[1], self._org_c[2]])
        return (origin, pos, neg, rough, vis)

    def combine_maps_async(self):
        """Extension: combine_maps that returns at once with a PendingMaps handle; the scan
        pipeline can be fed while the maps are still being produced / copied.  `.result()` gives the
        5-tuple of combine_maps (numpy views of a fresh pinned block), or None if the buffer was empty."""
        pos, neg, rough, vis = self._out_arrays()
        rc = check(self._L.gvom_combine_maps_async(self._h, self._org_c, pos.ctypes.data, neg.ctypes.data,
                                                   rough.ctypes.data, vis.ctypes.data, GVOM_HOST, self._stream),
                   "gvom_combine_maps_async")
        if rc == GVOM_NO_DATA:
            print("ERROR: No data in buffer")
            return PendingMaps(self, None)
        origin = np.array([self._org_c[0], self._org_c[1], self._org_c[2]])
        return PendingMaps(self, (origin, pos, neg, rough, vis))

    def _grid_block(self, device_outputs):
        S = self.xy_size
        if device_outputs:
            t = self._torch.empty((len(GRID_NAMES), S * S), dtype=self._torch.int8, device=f"cuda:{self.device}")
            return t, t.data_ptr(), GVOM_DEVICE
        if self.pinned_outputs:
            a = self._torch.empty((len(GRID_NAMES), S * S), dtype=self._torch.int8, pin_memory=True).numpy()
        else:
            a = np.empty((len(GRID_NAMES), S * S), np.int8)
        return a, a.ctypes.data, GVOM_HOST

    def occupancy_grids(self, density_threshold=50, min_roughness=-10, max_roughness=0, device_outputs=False):
        """Extension: the OccupancyGrid payloads the reference node builds on the host from
        the combine_maps outputs (gvom_ros.py:142-164), computed on the device from the last combine:
        dict of int8 arrays of xy_size*xy_size cells flattened in Fortran order ('hard', 'soft',
        'certainty', 'negative', 'roughness').  None before the first combine."""
        blk, ptr, mem = self._grid_block(device_outputs)
        rc = check(self._L.gvom_occupancy_grids(self._h, float(density_threshold), float(min_roughness),
                                                float(max_roughness), ptr, mem, self._stream), "gvom_occupancy_grids")
        if rc == GVOM_NO_DATA:
            print("No data")
            return None
        return {k: blk[i] for i, k in enumerate(GRID_NAMES)}

    def combine_maps_grids(self, density_threshold=50, min_roughness=-10, max_roughness=0, device_outputs=False):
        """combine_maps() + occupancy_grids() in one call; only the int8 grids leave the device.
        Returns (origin, grids dict) or None when the buffer is empty."""
        blk, ptr, mem = self._grid_block(device_outputs)
        rc = check(self._L.gvom_combine_maps_grids(self._h, self._org_c, float(density_threshold), float(min_roughness),
                                                   float(max_roughness), ptr, mem, self._stream), "gvom_combine_maps_grids")
        if rc == GVOM_NO_DATA:
            print("ERROR: No data in buffer")
            return None
        origin = np.array([self._org_c[0], self._org_c[1], self._org_c[2]])
        return origin, {k: blk[i] for i, k in enumerate(GRID_NAMES)}

    # ------------------------------------------------------------------ state
    def save_state(self, path=None):
        """Extension: ring slots + last combined map + ego as one opaque uint8 array
        (written to `path` with numpy.save if given)."""
        n, w = C.c_size_t(0), C.c_size_t(0)
        for _ in range(8):
            check(self._L.gvom_state_size(self._h, C.byref(n)), "gvom_state_size")
            blob = np.empty(n.value, np.uint8)
            rc = self._L.gvom_save_state(self._h, blob.ctypes.data, blob.size, C.byref(w))
            if rc != 3:                  # GVOM_ECAPACITY: another thread processed a scan between the two calls
                break
        check(rc, "gvom_save_state")
        blob = blob[:w.value]
        if path is not None:
            np.save(path, blob, allow_pickle=False)
        return blob

    def load_state(self, blob):
        """Restore a state saved by save_state (array or path) into this Gvom (same parameters)."""
        if isinstance(blob, str):
            blob = np.load(blob, allow_pickle=False)
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        check(self._L.gvom_load_state(self._h, blob.ctypes.data, blob.size), "gvom_load_state")

    # ------------------------------------------------------------------ debug
    def make_debug_voxel_map(self):
        n = C.c_int64(0)
        if self._L.gvom_combined_cell_count(self._h, C.byref(n)) == GVOM_NO_DATA:
            print("No data")
            return None
        out = np.zeros((n.value, 8), np.float32)
        rows = C.c_int64(0)
        check(self._L.gvom_debug_voxel_map(self._h, out.ctypes.data, n.value, C.byref(rows)), "gvom_debug_voxel_map")
        return out

    def make_debug_height_map(self):
        out = np.zeros((self.xy_size * self.xy_size, 7), np.float32)
        if check(self._L.gvom_debug_height_map(self._h, out.ctypes.data), "gvom_debug_height_map") == GVOM_NO_DATA:
            print("No data")
            return None
        return out

    def make_debug_inferred_height_map(self):
        out = np.zeros((self.xy_size * self.xy_size, 3), np.float32)
        if check(self._L.gvom_debug_inferred_height_map(self._h, out.ctypes.data),
                 "gvom_debug_inferred_height_map") == GVOM_NO_DATA:
            print("No data")
            return None
        return out

    def synchronize(self):
        """Wait until everything enqueued on the Gvom's stream (scans, asynchronous combines) has finished."""
        self._work_stream.synchronize()

    # ---------------------------------------------------------------- tooling
    def stats(self):
        s = GvomStats()
        check(self._L.gvom_get_stats(self._h, C.byref(s)), "gvom_get_stats")
        return {k: getattr(s, k) for k, _ in GvomStats._fields_}

    def set_variant(self, mask):
        """Tooling: select earlier kernel builds (bit mask, gvom_api.cu VAR_*); 0 = current."""
        check(self._L.gvom_set_variant(self._h, int(mask)), "gvom_set_variant")

    def set_profiling(self, on):
        check(self._L.gvom_set_profiling(self._h, int(bool(on))), "gvom_set_profiling")

    def stage_times(self):
        ms = (C.c_float * 16)()
        check(self._L.gvom_stage_times(self._h, ms), "gvom_stage_times")
        names = ("h2d", "scan_points", "scan_cells", "_3", "_4", "merge_codes", "merge_cells", "maps", "d2h", "stage_copy_host", "_10", "_11", "_12", "push_or_partial", "_14", "rows_surface")
        return {k: float(ms[i]) for i, k in enumerate(names)}

    def refview(self):
        """Host copy of the state under the reference's attribute names (test hook)."""
        return _RefView(self)


class PendingMaps:
    """Outputs of combine_maps_async: `.result()` waits and returns the combine_maps 5-tuple (or None)."""

    def __init__(self, g, out):
        self._g, self._out, self._done = g, out, out is None

    def done(self):
        return self._done

    def result(self):
        if not self._done:
            check(self._g._L.gvom_combine_wait(self._g._h), "gvom_combine_wait")
            self._done = True
        return self._out


class _RefView:
    """Snapshot of a Gvom's device state named like the reference's attributes
    (gvom.py:50-70), for the canonical parity dumps in tests/canon.py."""

    def __init__(self, g):
        L, h = g._L, g._h
        self._g = g
        B, V, S = g.buffer_size, g.voxel_count, g.xy_size
        slot = C.c_int32(0)
        check(L.gvom_last_slot(h, C.byref(slot)), "gvom_last_slot")
        self.last_buffer_index = slot.value
        self.index_buffer, self.hit_count_buffer, self.total_count_buffer = [None] * B, [None] * B, [None] * B
        self.metrics_buffer, self.min_height_buffer, self.origin_buffer = [None] * B, [None] * B, [None] * B
        for i in range(B):
            valid, cells, org = C.c_int32(0), C.c_int64(0), (C.c_double * 3)()
            check(L.gvom_slot_info(h, i, C.byref(valid), C.byref(cells), org), "gvom_slot_info")
            if not valid.value:
                continue
            n = cells.value
            idx, hit, tot = np.empty(V, np.int32), np.empty(n, np.int32), np.empty(n, np.int32)
            met, mnh = np.empty((n, 10), np.float64), np.empty(n, np.float32)
            check(L.gvom_export_slot(h, i, idx.ctypes.data, hit.ctypes.data, tot.ctypes.data, met.ctypes.data,
                                     mnh.ctypes.data), "gvom_export_slot")
            self.index_buffer[i], self.hit_count_buffer[i], self.total_count_buffer[i] = idx, hit, tot
            self.metrics_buffer[i], self.min_height_buffer[i] = met, mnh
            self.origin_buffer[i] = np.array(list(org))
        n = C.c_int64(0)
        self.combined_index_map = None
        if L.gvom_combined_cell_count(h, C.byref(n)) != GVOM_NO_DATA:
            n = n.value
            self.combined_cell_count_cpu = n
            self.combined_index_map = np.empty(V, np.int32)
            self.combined_hit_count, self.combined_total_count = np.empty(n, np.int32), np.empty(n, np.int32)
            self.combined_min_height = np.empty(n, np.float32)
            self.combined_metrics, self.voxels_eigenvalues = np.empty((n, 10), np.float32), np.empty((n, 3), np.float32)
            maps = np.empty((6, S, S), np.float64)
            check(L.gvom_export_combined(h, self.combined_index_map.ctypes.data, self.combined_hit_count.ctypes.data,
                                         self.combined_total_count.ctypes.data, self.combined_min_height.ctypes.data,
                                         self.combined_metrics.ctypes.data, self.voxels_eigenvalues.ctypes.data,
                                         maps.ctypes.data), "gvom_export_combined")
            (self.height_map, self.inferred_height_map, self.roughness_map, self.x_slope_map, self.y_slope_map,
             self.guessed_height_delta) = maps

    def make_debug_voxel_map(self):
        return self._g.make_debug_voxel_map()

    def make_debug_height_map(self):
        return self._g.make_debug_height_map()

    def make_debug_inferred_height_map(self):
        return self._g.make_debug_inferred_height_map()
