"""ctypes binding of the C-ABI in include/gvom_b200.h (libgvom_b200.so).

There is deliberately NO fallback: if the CUDA library is missing or cannot be
loaded, importing this module's `lib()` raises.  The CPU oracle under oracle/ is
test infrastructure and is never reachable from here.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libgvom_b200.so")

GVOM_OK, GVOM_NO_DATA = 0, -1
GVOM_F32, GVOM_F64 = 0, 1
GVOM_HOST, GVOM_DEVICE, GVOM_NONE = 0, 1, 2
GRID_NAMES = ("hard", "soft", "certainty", "negative", "roughness")     # GVOM_GRID_* order
RECORD_FLOATS = 16


class GvomParams(C.Structure):
    _fields_ = [("xy_resolution", C.c_double), ("z_resolution", C.c_double),
                ("xy_size", C.c_int32), ("z_size", C.c_int32), ("buffer_size", C.c_int32),
                ("_pad0", C.c_int32),
                ("min_distance", C.c_double), ("positive_obstacle_threshold", C.c_double),
                ("negative_obstacle_threshold", C.c_double), ("slope_obsacle_threshold", C.c_double),
                ("robot_height", C.c_double), ("robot_radius", C.c_double),
                ("ground_to_lidar_height", C.c_double),
                ("xy_eigen_dist", C.c_int32), ("z_eigen_dist", C.c_int32)]


class GvomStats(C.Structure):
    _fields_ = [("scan_cells", C.c_int64), ("combined_cells", C.c_int64),
                ("kernel_launches", C.c_int64), ("process_calls", C.c_int64),
                ("combine_calls", C.c_int64)]


MAX_RANKS = 16


class GvomRowsLinks(C.Structure):
    """include/gvom_b200.h: GvomRowsLinks (2-D exchange of the mirrored, row-sharded multi-GPU combine)."""
    _fields_ = [("rank", C.c_int32), ("nranks", C.c_int32),
                ("blocks2d", C.c_void_p * MAX_RANKS),
                ("heights_slots", C.c_void_p * MAX_RANKS), ("heights_flags", C.c_void_p),
                ("results_slots", C.c_void_p * MAX_RANKS), ("results_flags", C.c_void_p)]


# every symbol include/gvom_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
_pd = C.POINTER(C.c_double)
SYMBOLS = {
    "gvom_last_error": (C.c_char_p, []),
    "gvom_workspace_size": (C.c_int, [C.POINTER(GvomParams), _i64, _i64, C.POINTER(_sz), C.POINTER(_sz)]),
    "gvom_create": (C.c_int, [C.POINTER(GvomParams), _i64, _i64, C.c_int, _vp, _sz, _vp, _sz, C.POINTER(_vp)]),
    "gvom_destroy": (C.c_int, [_vp]),
    "gvom_process_pointcloud": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _pd, _vp, _vp]),
    "gvom_get_stream": (C.c_int, [_vp, C.POINTER(_vp)]),
    "gvom_wait_input": (C.c_int, [_vp]),
    "gvom_combine_maps": (C.c_int, [_vp, _pd, _vp, _vp, _vp, _vp, _i32, _vp]),
    "gvom_process_pointcloud2": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _pd, _vp, _vp]),
    "gvom_combine_maps_async": (C.c_int, [_vp, _pd, _vp, _vp, _vp, _vp, _i32, _vp]),
    "gvom_combine_wait": (C.c_int, [_vp]),
    "gvom_occupancy_grids": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double, _vp, _i32, _vp]),
    "gvom_combine_maps_grids": (C.c_int, [_vp, _pd, C.c_double, C.c_double, C.c_double, _vp, _i32, _vp]),
    "gvom_state_size": (C.c_int, [_vp, C.POINTER(_sz)]),
    "gvom_save_state": (C.c_int, [_vp, _vp, _sz, C.POINTER(_sz)]),
    "gvom_load_state": (C.c_int, [_vp, _vp, _sz]),
    "gvom_set_variant": (C.c_int, [_vp, C.c_uint32]),
    "gvom_combined_cell_count": (C.c_int, [_vp, C.POINTER(_i64)]),
    "gvom_debug_voxel_map": (C.c_int, [_vp, _vp, _i64, C.POINTER(_i64)]),
    "gvom_debug_height_map": (C.c_int, [_vp, _vp]),
    "gvom_debug_inferred_height_map": (C.c_int, [_vp, _vp]),
    "gvom_newest_origin": (C.c_int, [_vp, _pd]),
    "gvom_combine_partial": (C.c_int, [_vp, _pd, _vp, _vp, _vp, _i64, _vp, C.POINTER(_vp), _i32, _i32, _vp]),
    "gvom_combine_finish": (C.c_int, [_vp, _pd, C.POINTER(_vp), C.POINTER(_vp), _i32, C.POINTER(_vp), C.POINTER(_vp), _i32, _i64, _vp, _i32, _pd,
                                      _vp, _vp, _vp, _vp, _i32, _vp]),
    "gvom_rows_block_size": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "gvom_combine_finish_rows": (C.c_int, [_vp, _pd, C.POINTER(GvomRowsLinks), _i32, _i32, _pd, _vp, _vp, _vp, _vp, _i32, _vp]),
    "gvom_mirror_block_size": (C.c_int, [_vp, _i32, C.POINTER(C.c_uint64)]),
    "gvom_mirror_attach": (C.c_int, [_vp, _i32, _i32, C.POINTER(_vp)]),
    "gvom_adopt_ego": (C.c_int, [_vp, _pd]),
    "gvom_slot_info": (C.c_int, [_vp, _i32, C.POINTER(_i32), C.POINTER(_i64), _pd]),
    "gvom_last_slot": (C.c_int, [_vp, C.POINTER(_i32)]),
    "gvom_export_slot": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "gvom_export_combined": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gvom_get_stats": (C.c_int, [_vp, C.POINTER(GvomStats)]),
    "gvom_set_profiling": (C.c_int, [_vp, _i32]),
    "gvom_stage_times": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "gvom_graph_probe": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _pd, _vp, _i32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(_i32)]),
    "gvom_bench_atomics": (C.c_int, [C.c_int, _vp, _i64, _i32, _i32, _i32, C.POINTER(C.c_float), C.POINTER(_i64)]),
}

_LIB = None


def lib():
    """Load libgvom_b200.so and bind every symbol of the header.  Raises if absent."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} not found: build the CUDA library first "
                "(python -c 'import __graft_entry__ as g; g.build()' or make -C gvom_b200/csrc). "
                "gvom_b200 has no CPU fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)            # AttributeError = header / library mismatch
            f.restype, f.argtypes = res, args
        _LIB = L
    return _LIB


def check(rc, what):
    if rc > 0:
        raise RuntimeError(f"{what} failed (code {rc}): {lib().gvom_last_error().decode()}")
    return rc
