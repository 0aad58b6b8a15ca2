"""ctypes access to the C restatement (oracle/gather_ref.c).  CPU ORACLE — test
infrastructure only: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never from the product package."""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libadell_ref.so")
_lib = None


def build() -> str:
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        lib.adell_ref_gather.restype = C.c_int
        lib.adell_ref_gather.argtypes = [C.c_void_p, C.c_int]
        lib.adell_ref_key_f32.restype = C.c_uint32
        lib.adell_ref_key_f32.argtypes = [C.c_float]
        lib.adell_ref_minmax.restype = None
        lib.adell_ref_minmax.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        lib.adell_ref_intensity_map.restype = None
        lib.adell_ref_intensity_map.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]
        lib.adell_ref_percentile_lerp.restype = C.c_float
        lib.adell_ref_percentile_lerp.argtypes = [C.c_float, C.c_float, C.c_double]
        lib.adell_ref_order_stats.restype = C.c_int
        lib.adell_ref_order_stats.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
        _lib = lib
    return _lib


def gather(items: np.ndarray) -> None:
    """Run the scalar C restatement over an ``adell_item`` array holding HOST pointers."""
    st = load().adell_ref_gather(items.ctypes.data, items.shape[0])
    if st != 0:
        raise RuntimeError(f"adell_ref_gather status {st}")


def intensity_map(x: np.ndarray, coefs, clip=False, lo=0.0, hi=0.0) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    c = np.ascontiguousarray(coefs, np.float32)
    y = np.empty_like(x)
    load().adell_ref_intensity_map(x.ctypes.data, x.size, c.ctypes.data, int(clip), lo, hi, y.ctypes.data)
    return y


def percentile_lerp(a: float, b: float, t: float) -> float:
    return float(load().adell_ref_percentile_lerp(a, b, t))
