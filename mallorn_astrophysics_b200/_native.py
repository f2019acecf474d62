"""ctypes binding of ``libmallorn_b200.so`` (C-ABI in ``include/mallorn_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``build_native.py`` with
``nvcc -gencode arch=compute_100a,code=sm_100a``.  There is deliberately no fallback: a missing
library or a machine without a CUDA device raises ``MallornNativeError`` on first use.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MLRN_LIB") or os.path.join(HERE, "libmallorn_b200.so")   # MLRN_LIB: developer builds (profiling flags)

COLS = {"statistical": 123, "colors": 83, "shape": 65, "tde": 25, "physics": 32,
        "bazin": 52, "powerlaw": 27, "gp2d": 27, "gp1d": 21, "research": 40}
SEGMETA_COLS = 22
N_SEG = 7
MAX_OBS = 65535       # within-object ranks are uint16; objects above SHORT_LIMIT (512) are staged in a global scratch


class MallornNativeError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load the CUDA library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MallornNativeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  This package has no CPU implementation to fall back to.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.mlrn_last_error.restype = C.c_char_p
    L.mlrn_version.restype = C.c_int
    L.mlrn_device_count.restype = C.c_int
    L.mlrn_launch_count.restype = C.c_longlong
    L.mlrn_measure_fp64.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.mlrn_store_create.argtypes = [C.c_int, i64, i32, vp, vp, vp, vp, vp, C.c_int, vp, C.POINTER(vp)]
    L.mlrn_store_destroy.argtypes = [vp]
    L.mlrn_store_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]
    L.mlrn_store_segments.argtypes = [vp, vp]
    L.mlrn_store_export.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    for name in ("mlrn_stat", "mlrn_color", "mlrn_shape", "mlrn_tde"):
        if hasattr(L, name):
            getattr(L, name).argtypes = [vp, vp, C.c_int, vp]
    for name in ("mlrn_physics", "mlrn_research"):
        if hasattr(L, name):
            getattr(L, name).argtypes = [vp, vp, C.c_int, vp, C.c_int, vp]
    for name in ("mlrn_bazin", "mlrn_powerlaw", "mlrn_gp2d", "mlrn_gp1d"):
        if hasattr(L, name):
            getattr(L, name).argtypes = [vp, vp, vp, C.c_int, vp]
    if hasattr(L, "mlrn_fits"):
        L.mlrn_fits.argtypes = [vp, vp, vp, vp, vp, C.c_int, vp]
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().mlrn_last_error()
        raise MallornNativeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def require_device() -> int:
    n = lib().mlrn_device_count()
    if n <= 0:
        raise MallornNativeError("no CUDA device visible: mallorn_astrophysics_b200 computes on the GPU only "
                                 "(there is no CPU fallback)")
    return n
