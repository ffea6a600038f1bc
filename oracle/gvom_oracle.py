"""Python driver of the CPU parity oracle (oracle/gvom_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by gvom_b200.

`OracleGvom` restates the HOST side of the reference class (gvom.py:21-442,
1069-1119): ring buffer of per-scan maps, call order of the kernels, state
carried between combine_maps() calls.  It keeps the reference's attribute
names (index_buffer, combined_index_map, ...) as plain numpy arrays so the
same canonicalisation code can dump the reference, the oracle and the CUDA
path.  The kernels themselves are the C functions, one per reference kernel.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Params(C.Structure):
    _fields_ = [("xy_res", C.c_double), ("z_res", C.c_double),
                ("xy_size", C.c_int64), ("z_size", C.c_int64),
                ("min_distance", C.c_double),
                ("pos_thr", C.c_double), ("neg_thr", C.c_double), ("slope_thr", C.c_double),
                ("robot_height", C.c_double), ("robot_radius", C.c_double),
                ("ground_to_lidar", C.c_double),
                ("xy_eigen_dist", C.c_int64), ("z_eigen_dist", C.c_int64)]


def build(force=False):
    so = os.path.join(HERE, "libgvom_oracle.so")
    src = os.path.join(HERE, "gvom_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "libgvom_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.gvo_point_2_map.restype = C.c_int64
        _LIB.gvo_assign_indices.restype = C.c_int64
        _LIB.gvo_moments.restype = C.c_int64
        _LIB.gvo_max_threads.restype = C.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleGvom:
    """CPU restatement of reference `Gvom` (gvom.py:8-442)."""

    def __init__(self, xy_resolution, z_resolution, xy_size, z_size, buffer_size, min_distance,
                 positive_obstacle_threshold, negative_obstacle_threshold, slope_obsacle_threshold,
                 robot_height, robot_radius, ground_to_lidar_height, xy_eigen_dist, z_eigen_dist):
        self.L = lib()
        self.xy_resolution, self.z_resolution = xy_resolution, z_resolution
        self.xy_size, self.z_size, self.buffer_size = xy_size, z_size, buffer_size
        self.robot_radius, self.ground_to_lidar_height = robot_radius, ground_to_lidar_height
        self.voxel_count = xy_size * xy_size * z_size
        self.P = _Params(xy_resolution, z_resolution, xy_size, z_size, min_distance,
                         positive_obstacle_threshold, negative_obstacle_threshold,
                         slope_obsacle_threshold, robot_height, robot_radius,
                         ground_to_lidar_height, xy_eigen_dist, z_eigen_dist)
        B = buffer_size
        self.index_buffer = [None] * B
        self.hit_count_buffer = [None] * B
        self.total_count_buffer = [None] * B
        self.metrics_buffer = [None] * B
        self.origin_buffer = [None] * B
        self.min_height_buffer = [None] * B
        self.buffer_index = 0
        self.last_buffer_index = 0
        self.ego_position = [0, 0, 0]
        self.combined_cell_count_cpu = None
        self.combined_index_map = None
        self.last_combined_origin = None
        self.height_map = None
        self.work = {}

    # gvom.py:105-220
    def Process_pointcloud(self, pointcloud, ego_position, transform=None):
        L, P = self.L, C.byref(self.P)
        self.ego_position = ego_position
        pc = np.array(pointcloud, copy=True, order="C")       # cuda.to_device keeps the dtype
        if pc.dtype not in (np.float32, np.float64):
            pc = pc.astype(np.float64)
        n, stride, f32 = pc.shape[0], pc.shape[1], int(pc.dtype == np.float32)
        V = self.voxel_count
        origin = np.zeros(3)
        origin[0] = math.floor((ego_position[0] / self.xy_resolution) - self.xy_size / 2)
        origin[1] = math.floor((ego_position[1] / self.xy_resolution) - self.xy_size / 2)
        origin[2] = math.floor((ego_position[2] / self.z_resolution) - self.z_size / 2)
        ego = np.asarray(ego_position, dtype=np.float64)
        if transform is not None:
            T = np.ascontiguousarray(transform, dtype=np.float64)
            L.gvo_transform(_p(pc), f32, C.c_int64(stride), C.c_int64(n), _p(T))
        hit = np.zeros(V, np.int32)
        tot = np.zeros(V, np.int32)
        steps = L.gvo_point_2_map(P, _p(pc), f32, C.c_int64(stride), C.c_int64(n), _p(ego), _p(origin),
                                  _p(hit), _p(tot))
        index_map = np.empty(V, np.int32)
        cap = int(min(n, V))
        hit_c = np.empty(cap, np.int32)
        tot_c = np.empty(cap, np.int32)
        cells = int(L.gvo_assign_indices(C.c_int64(V), _p(hit), _p(tot), _p(index_map), _p(hit_c), _p(tot_c)))
        hit_c, tot_c = hit_c[:cells].copy(), tot_c[:cells].copy()
        metrics = np.zeros((cells, 10), np.float64)
        K = L.gvo_moments(P, _p(pc), f32, C.c_int64(stride), C.c_int64(n), _p(origin), _p(index_map),
                          C.c_int64(cells), _p(metrics))
        min_height = np.ones(cells * 3, np.float32)              # over-allocated x3 (gvom.py:1084)
        L.gvo_min_height(P, _p(pc), f32, C.c_int64(stride), C.c_int64(n), _p(origin), _p(index_map),
                         C.c_int64(cells), _p(min_height))
        self.work = {"n": n, "n_in": int(hit_c.sum()), "cells": cells, "dda_steps": int(steps), "pairs": int(K)}
        i = self.buffer_index
        self.index_buffer[i], self.hit_count_buffer[i], self.total_count_buffer[i] = index_map, hit_c, tot_c
        self.metrics_buffer[i], self.min_height_buffer[i], self.origin_buffer[i] = metrics, min_height, origin
        self.last_buffer_index = i
        self.buffer_index = (i + 1) % self.buffer_size

    process_pointcloud = Process_pointcloud

    # gvom.py:222-393
    def combine_maps(self):
        L, P = self.L, C.byref(self.P)
        if self.origin_buffer[self.last_buffer_index] is None:
            print("ERROR: No data in buffer")
            return None
        V, S = self.voxel_count, self.xy_size
        self.combined_origin = self.origin_buffer[self.last_buffer_index].copy()
        co = self.combined_origin
        counter = np.zeros(1, np.int64)
        cmap = np.full(V, -1, np.int32)
        for i in range(self.buffer_size):
            if self.origin_buffer[i] is None:
                continue
            L.gvo_combine_indices(P, _p(counter), _p(cmap), _p(co), _p(self.index_buffer[i]),
                                  _p(self.origin_buffer[i]), 0)
        if self.last_combined_origin is not None:
            L.gvo_combine_indices(P, _p(counter), _p(cmap), _p(co), _p(self.last_combined_index_map),
                                  _p(self.last_combined_origin), 1)
        Cc = int(counter[0])
        chit = np.zeros(Cc, np.int32)
        ctot = np.zeros(Cc, np.int32)
        cminh = np.ones(Cc, np.float32)
        cm = np.zeros((Cc, 10), np.float32)
        for i in range(self.buffer_size):
            if self.origin_buffer[i] is None:
                continue
            L.gvo_combine_metrics(P, _p(cm), _p(chit), _p(ctot), _p(cminh), _p(cmap), _p(co),
                                  _p(self.metrics_buffer[i]), 0, _p(self.hit_count_buffer[i]),
                                  _p(self.total_count_buffer[i]), _p(self.min_height_buffer[i]),
                                  _p(self.index_buffer[i]), _p(self.origin_buffer[i]))
        if self.last_combined_origin is not None:
            L.gvo_combine_metrics(P, _p(cm), _p(chit), _p(ctot), _p(cminh), _p(cmap), _p(co),
                                  _p(self.last_combined_metrics), 1, _p(self.last_combined_hit_count),
                                  _p(self.last_combined_total_count), _p(self.last_combined_min_height),
                                  _p(self.last_combined_index_map), _p(self.last_combined_origin))
        self.combined_cell_count_cpu = Cc
        self.combined_index_map, self.combined_hit_count, self.combined_total_count = cmap, chit, ctot
        self.combined_min_height, self.combined_metrics = cminh, cm
        self.last_combined_index_map, self.last_combined_hit_count = cmap, chit
        self.last_combined_total_count, self.last_combined_min_height = ctot, cminh
        self.last_combined_metrics, self.last_combined_origin = cm, co

        self.voxels_eigenvalues = np.zeros((Cc, 3), np.float32)
        L.gvo_eigenvalues(C.c_int64(Cc), _p(cm), _p(self.voxels_eigenvalues))

        ego = np.asarray(self.ego_position, dtype=np.float64)
        self.height_map = np.empty((S, S))
        self.inferred_height_map = np.empty((S, S))
        L.gvo_height_maps(P, _p(co), _p(cmap), _p(cminh), _p(ego), _p(self.height_map),
                          _p(self.inferred_height_map))
        self.roughness_map = np.empty((S, S))
        self.x_slope_map = np.empty((S, S))
        self.y_slope_map = np.empty((S, S))
        L.gvo_slope(P, _p(self.height_map), _p(self.x_slope_map), _p(self.y_slope_map), _p(self.roughness_map))
        self.guessed_height_delta = np.empty((S, S))
        L.gvo_guess_height(P, _p(self.height_map), _p(self.inferred_height_map), _p(self.guessed_height_delta))
        pos = np.empty((S, S), np.int32)
        neg = np.empty((S, S), np.int32)
        vis = np.empty((S, S), np.int32)
        L.gvo_obstacle_maps(P, _p(co), _p(cmap), _p(chit), _p(ctot), _p(self.height_map), _p(self.x_slope_map),
                            _p(self.y_slope_map), _p(self.guessed_height_delta), _p(pos), _p(neg), _p(vis))
        ow = co.copy()
        ow[0] *= self.xy_resolution
        ow[1] *= self.xy_resolution
        ow[2] *= self.z_resolution
        return (ow, pos, neg, self.roughness_map.copy(), vis)

    # gvom.py:395-442
    def make_debug_voxel_map(self):
        if self.combined_cell_count_cpu is None:
            print("No data")
            return None
        out = np.zeros((self.combined_cell_count_cpu, 8), np.float32)
        self.L.gvo_debug_voxel_map(C.byref(self.P), _p(self.combined_origin), _p(self.combined_index_map),
                                   _p(self.combined_hit_count), _p(self.combined_total_count),
                                   _p(self.voxels_eigenvalues), _p(out))
        return out

    def make_debug_height_map(self):
        if self.height_map is None:
            print("No data")
            return None
        out = np.zeros((self.xy_size * self.xy_size, 7), np.float32)
        self.L.gvo_debug_height_map(C.byref(self.P), _p(self.combined_origin), _p(self.height_map),
                                    _p(self.roughness_map), _p(self.x_slope_map), _p(self.y_slope_map), _p(out))
        return out

    def make_debug_inferred_height_map(self):
        if self.height_map is None:
            print("No data")
            return None
        out = np.zeros((self.xy_size * self.xy_size, 3), np.float32)
        self.L.gvo_debug_inferred_height_map(C.byref(self.P), _p(self.combined_origin),
                                             _p(self.guessed_height_delta), _p(out))
        return out
