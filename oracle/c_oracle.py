"""ctypes loader of oracle/liboracle_pt.so (the C restatement, pt_oracle.c).
TEST INFRASTRUCTURE ONLY -- see the header of pt_oracle.c."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle_pt.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        lib.oracle_triples_list.restype = C.c_int
        lib.oracle_triples_list.argtypes = [C.c_int, C.c_int] + [dp] * 7 + [C.POINTER(C.c_int64), C.c_int64, dp, C.c_int]
        lib.oracle_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def max_threads() -> int:
    return int(load().oracle_max_threads())


def triples_list(epsi, epsa, T1, T2, Vpphh, Vhhhp, Vppph, idx, nthreads: int = 0) -> np.ndarray:
    """E_t of the sorted triples idx (reference enumeration order), C restatement."""
    lib = load()
    o, v = int(epsi.size), int(epsa.size)
    arrs = [np.asfortranarray(a, dtype=np.float64) for a in (epsi, epsa, T1, T2, Vpphh, Vhhhp, Vppph)]
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    out = np.zeros(idx.size, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    rc = lib.oracle_triples_list(o, v, *[a.ctypes.data_as(dp) for a in arrs],
                                 idx.ctypes.data_as(C.POINTER(C.c_int64)), idx.size,
                                 out.ctypes.data_as(dp), int(nthreads))
    if rc != 0:
        raise MemoryError("oracle_triples_list: allocation failed")
    return out
