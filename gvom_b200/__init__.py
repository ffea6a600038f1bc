"""gvom_b200 -- B200-native implementation of G-VOM's per-scan voxel-mapping path.

`from gvom_b200 import Gvom` gives the drop-in class (same constructor and
methods as the reference scripts/gvom.py); adding gvom_b200/shim to PYTHONPATH
makes `import gvom` resolve to it, so the reference's gvom_ros.py runs unchanged.
The class is a thin host over the C-ABI in include/gvom_b200.h.
"""


def __getattr__(name):
    if name in ("Gvom", "MultiGpuGvom"):
        from . import gvom as _g
        return getattr(_g, name)
    raise AttributeError(name)
