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
        pp = C.POINTER(C.c_void_p)
        lib.oracle_triples_list_blocks.restype = C.c_int
        lib.oracle_triples_list_blocks.argtypes = [C.c_int, C.c_int] + [dp] * 4 + [pp, dp, pp] + [C.POINTER(C.c_int64), C.c_int64, dp, C.c_int]
        lib.oracle_max_threads.restype = C.c_int
        lib.oracle_set_dgemm.argtypes = [C.c_void_p]
        lib.oracle_has_dgemm.restype = C.c_int
        _lib = lib
    return _lib


def use_blas(on: bool = True) -> bool:
    """Route the two GEMMs of getDoublesContribution through the dgemm of the OpenBLAS bundled with
    scipy (the reference reaches the system dgemm_ through CTF).  Returns whether a BLAS is in use."""
    lib = load()
    if not on:
        lib.oracle_set_dgemm(None)
        return False
    try:
        from scipy.linalg import cython_blas
        cap = cython_blas.__pyx_capi__["dgemm"]
        get = C.pythonapi.PyCapsule_GetPointer
        get.restype = C.c_void_p
        get.argtypes = [C.py_object, C.c_char_p]
        name = C.pythonapi.PyCapsule_GetName
        name.restype = C.c_char_p
        name.argtypes = [C.py_object]
        lib.oracle_set_dgemm(get(cap, name(cap)))
    except Exception:
        lib.oracle_set_dgemm(None)
    return bool(lib.oracle_has_dgemm())


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


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


def triple_of(o: int, t: int):
    """Sorted triple number t of the reference enumeration (CcsdPerturbativeTriples.cxx:156-158)."""
    n = 0
    for i in range(o):
        for j in range(i, o):
            if t < n + (o - j):
                return i, j, j + (t - n)
            n += o - j
    raise IndexError(t)


def triples_list_blocks(epsi, epsa, T1, T2, pphh_block, Vhhhp, ppph_slab, idx, nthreads: int = 0) -> np.ndarray:
    """E_t of the sorted triples idx with the big integral tensors given lazily: ``ppph_slab(z)`` returns
    Vppph[:,:,:,z] ([v,v,v], column-major) and ``pphh_block(j, k)`` returns Vpphh[:,:,j,k] ([v,v]); only
    the slabs / blocks the listed triples touch are asked for (o=64 v=512 and o=100 v=800 checks)."""
    lib = load()
    o, v = int(epsi.size), int(epsa.size)
    arrs = [np.asfortranarray(a, dtype=np.float64) for a in (epsi, epsa, T1, T2, Vhhhp)]
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    holes = sorted({h for t in idx for h in triple_of(o, int(t))})
    keep = []
    slabs = (C.c_void_p * o)()
    pairs = (C.c_void_p * (o * o))()
    for z in holes:
        a = np.asfortranarray(ppph_slab(z), dtype=np.float64)
        keep.append(a)
        slabs[z] = a.ctypes.data
    for j in holes:
        for k in holes:
            a = np.asfortranarray(pphh_block(j, k), dtype=np.float64)
            keep.append(a)
            pairs[j + o * k] = a.ctypes.data
    out = np.zeros(idx.size, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    rc = lib.oracle_triples_list_blocks(o, v, *[a.ctypes.data_as(dp) for a in arrs[:4]], pairs, arrs[4].ctypes.data_as(dp),
                                        slabs, idx.ctypes.data_as(C.POINTER(C.c_int64)), idx.size,
                                        out.ctypes.data_as(dp), int(nthreads))
    if rc != 0:
        raise MemoryError("oracle_triples_list_blocks: allocation failed")
    return out
