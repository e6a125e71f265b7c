"""ctypes binding of libsisi4s_pt.so (include/sisi4s_pt.h).

The library is built in-tree by ``__graft_entry__.build()`` (or ``make -C
sisi4s_b200/csrc``).  There is no Python / CPU fallback: if the shared library
is missing or no CUDA device is present, loading or ``pt_create`` raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SISI4S_PT_LIB") or os.path.join(_HERE, "libsisi4s_pt.so")   # override: A/B builds

PT_ENGINE_FUSED = 0
PT_ENGINE_NAIVE = 1


class PtStats(C.Structure):
    _fields_ = [
        ("seconds_run", C.c_double),
        ("seconds_kernel", C.c_double),
        ("seconds_upload", C.c_double),
        ("flops_algorithmic", C.c_double),
        ("bytes_h2d", C.c_double),
        ("bytes_d2h", C.c_double),
        ("device_bytes", C.c_double),
        ("kernel_launches", C.c_int64),
        ("triples_run", C.c_int64),
        ("sm_count", C.c_int32),
        ("reserved", C.c_int32),
        ("slab_loads", C.c_int64),
        ("groups_staged", C.c_int64),
        ("bytes_pinned", C.c_double),
    ]


# every symbol include/sisi4s_pt.h declares: name -> (restype, argtypes)
_DP = C.POINTER(C.c_double)
_H = C.c_void_p
SYMBOLS = {
    "pt_create": (C.c_int, [C.POINTER(_H), C.c_int, C.c_int, C.c_int]),
    "pt_create_ex": (C.c_int, [C.POINTER(_H), C.c_int, C.c_int, C.c_int, C.c_int]),
    "pt_destroy": (C.c_int, [_H]),
    "pt_last_error": (C.c_char_p, []),
    "pt_version": (C.c_char_p, []),
    "pt_set_option": (C.c_int, [_H, C.c_char_p, C.c_int64]),
    "pt_sync": (C.c_int, [_H]),
    "pt_estimate_device_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "pt_use_vertex_integrals": (C.c_int, [_H]),
    "pt_vertex_integrals": (C.c_int, [_H, C.c_char_p, _DP]),
    "pt_set_eigenenergies": (C.c_int, [_H, _DP, _DP]),
    "pt_set_singles": (C.c_int, [_H, _DP]),
    "pt_set_singles_pair": (C.c_int, [_H, _DP, _DP]),
    "pt_set_doubles": (C.c_int, [_H, _DP]),
    "pt_set_doubles_hole": (C.c_int, [_H, _DP]),
    "pt_set_pphh": (C.c_int, [_H, _DP]),
    "pt_set_hhhp": (C.c_int, [_H, _DP]),
    "pt_set_ppph_slabs": (C.c_int, [_H, C.c_int, C.c_int, _DP]),
    "pt_set_ppph_host": (C.c_int, [_H, _DP]),
    "pt_set_vertex": (C.c_int, [_H, C.c_int, C.c_int, _DP, _DP]),
    "pt_num_triples": (C.c_int64, [C.c_int]),
    "pt_partition": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "pt_plan_hole_blocks": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int64)]),
    "pt_run": (C.c_int, [_H, C.c_int64, C.c_int64, _DP, _DP]),
    "pt_run_list": (C.c_int, [_H, C.c_int64, C.POINTER(C.c_int64), _DP, _DP]),
    "pt_get_stats": (C.c_int, [_H, C.POINTER(PtStats)]),
    "pt_spin_orbital_triples": (C.c_int, [C.c_int, C.c_int, C.c_int] + [_DP] * 8),
    "pt_complex_triples": (C.c_int, [C.c_int, C.c_int, C.c_int] + [_DP] * 10 + [C.c_int, C.c_int, _DP, _DP, _DP, _DP]),
    "pt_debug_w_tile": (C.c_int, [_H] + [C.c_int] * 6 + [_DP]),
    "pt_bench_fp64": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, _DP, _DP]),
    "pt_bench_vertex_gemm": (C.c_int, [_H, C.c_int, C.c_int, _DP, _DP]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library and bind every declared symbol (fails loudly)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C sisi4s_b200/csrc`.  There is no CPU fallback for the (T) step.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class PtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsisi4s_pt error {code}: {msg}")
        self.code = code


def check(rc: int) -> None:
    if rc != 0:
        raise PtError(rc, load().pt_last_error().decode())
