"""Print the launch geometry the library picks for the on-chip fit kernels."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from midagma_b200 import _lib
lib = _lib.load(); _lib.require_device()
for d in (16, 20, 32, 40, 48, 64):
    ctas, thr, sm = C.c_int(), C.c_int(), C.c_size_t()
    rc = lib.dagma_linear_fit_small_geometry(d, 100000, C.byref(ctas), C.byref(thr), C.byref(sm))
    print(d, rc, ctas.value, thr.value, sm.value, lib.dagma_last_error())
