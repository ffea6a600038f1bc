"""External shims that let the UNMODIFIED reference gvom.py run on numba 0.65.

Import this module BEFORE importing the reference `gvom`.  Nothing in the
reference source is edited; both shims restore behaviour of the numba releases
the reference was written against (2021, numba 0.5x):

1. Real GPU: the reference defines its kernels as `@cuda.jit` functions in the
   class body and launches them as `self.__kernel[grid, block](...)`
   (gvom.py:125 and every other launch).  In numba 0.65 CUDADispatcher inherits
   Dispatcher.__get__, which binds the kernel as a *method* (MethodType), and
   `method[grid, block]` raises "TypeError: 'method' object is not
   subscriptable" (measured on the B200 box, round 1).  Returning the
   dispatcher itself from __get__ restores the old non-binding behaviour.
2. CUDASIM: the reference spells `numba.cuda.local.array` through the `numba`
   global (gvom.py:734,1125,1174-1176,1464); the simulator only swaps globals
   that ARE the cuda module, so that attribute does not exist there.
"""
import os

import numba
import numba.cuda

CUDASIM = os.environ.get("NUMBA_ENABLE_CUDASIM") == "1"

if CUDASIM:
    from numba.cuda.simulator.kernelapi import FakeCUDALocal
    numba.cuda.local = FakeCUDALocal()
else:
    from numba.cuda.dispatcher import CUDADispatcher
    CUDADispatcher.__get__ = lambda self, obj, objtype=None: self
