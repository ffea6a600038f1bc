"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/gvom_b200.h declares, and the host class mirrors the reference's
public surface (scripts/gvom.py:21-22,105,222,395-442)."""
import ctypes as C
import inspect
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from gvom_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "gvom_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gvom_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations found"
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)


def test_header_is_plain_c_and_links(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 and a C program must link against the library."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "gvom_b200.h")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    src = tmp_path / "abi.c"
    src.write_text('''#include "gvom_b200.h"
#include <stdio.h>
int main(void) {
    GvomParams p = {0.4, 0.2, 256, 64, 4, 0, 1.0, 0.5, 0.5, 0.3, 2.0, 4.0, 1.0, 1, 1};
    size_t db = 0, hb = 0;
    int rc = gvom_workspace_size(&p, 262144, 0, &db, &hb);
    printf("%d %zu %zu %zu\\n", rc, db, hb, sizeof(GvomParams));
    p.buffer_size = 0;
    return rc != 0 || gvom_workspace_size(&p, 262144, 0, &db, &hb) != GVOM_EINVAL || gvom_last_error()[0] == 0;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.join(ROOT, "gvom_b200")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-lgvom_b200", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert out[0] == "0" and int(out[1]) > 1 << 28 and out[3] == "96"


def test_struct_layout_matches_header():
    from gvom_b200 import _lib
    assert C.sizeof(_lib.GvomParams) == 96
    assert _lib.GvomParams.min_distance.offset == 32
    assert _lib.GvomParams.xy_eigen_dist.offset == 88
    assert C.sizeof(_lib.GvomStats) == 40


def test_workspace_size_and_argument_errors_without_gpu():
    from gvom_b200 import _lib
    L = _lib.lib()
    P = _lib.GvomParams(0.4, 0.2, 256, 64, 4, 0, 1.0, 0.5, 0.5, 0.3, 2.0, 4.0, 1.0, 1, 1)
    db, hb = C.c_size_t(), C.c_size_t()
    assert L.gvom_workspace_size(C.byref(P), 1 << 18, 0, C.byref(db), C.byref(hb)) == 0
    assert db.value > 6 * 4 * 256 * 256 * 64 and hb.value > (1 << 18) * 24
    bad = _lib.GvomParams(0.4, 0.2, 256, 64, 0, 0, 1.0, 0.5, 0.5, 0.3, 2.0, 4.0, 1.0, 1, 1)
    assert L.gvom_workspace_size(C.byref(bad), 1 << 18, 0, C.byref(db), C.byref(hb)) == 1
    assert b"buffer_size" in L.gvom_last_error()


def test_class_surface_matches_reference():
    from gvom_b200.gvom import Gvom
    ref_args = ["xy_resolution", "z_resolution", "xy_size", "z_size", "buffer_size", "min_distance",
                "positive_obstacle_threshold", "negative_obstacle_threshold", "slope_obsacle_threshold",
                "robot_height", "robot_radius", "ground_to_lidar_height", "xy_eigen_dist", "z_eigen_dist"]
    sig = inspect.signature(Gvom.__init__)
    pos = [p.name for p in sig.parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD][1:]
    assert pos == ref_args
    assert list(inspect.signature(Gvom.Process_pointcloud).parameters)[1:] == ["pointcloud", "ego_position", "transform"]
    assert inspect.signature(Gvom.Process_pointcloud).parameters["transform"].default is None
    for m in ("combine_maps", "make_debug_voxel_map", "make_debug_height_map", "make_debug_inferred_height_map",
              "process_pointcloud"):
        assert callable(getattr(Gvom, m))


def test_no_cpu_fallback_and_oracle_not_imported_by_product():
    import torch
    from gvom_b200.gvom import Gvom
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            Gvom(0.4, 0.2, 32, 8, 2, 1.0, 0.5, 0.5, 0.3, 2.0, 4.0, 1.0, 1, 1)
    for root, _, files in os.walk(os.path.join(ROOT, "gvom_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "numba" not in src.replace("Numba", "") or f == "synth.py", f
