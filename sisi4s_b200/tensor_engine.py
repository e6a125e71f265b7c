"""ctypes binding of the device tensor-contraction engine (include/sisi4s_tn.h, csrc/tn_engine.cu).

One `DeviceTensors` object = one `tn_handle_t`: dense FP64 column-major tensors in the memory of one
GPU and the reference's CTF-style index-string statements on them,

    eng.contract(alpha, A, "acik", B, "cbkj", beta, C, "abij")     # C["abij"] = a A["acik"] B["cbkj"] + b C["abij"]
    eng.add(alpha, A, "abij", beta, C, "aibj")                     # C["aibj"] = a A["abij"] + b C["aibj"]

executed by the library's own FP64 tensor-core GEMM.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_DP = C.POINTER(C.c_double)
_H = C.c_void_p
TN_SYMBOLS = {
    "tn_create": (C.c_int, [C.POINTER(_H), C.c_int]),
    "tn_destroy": (C.c_int, [_H]),
    "tn_last_error": (C.c_char_p, []),
    "tn_tensor": (C.c_int, [_H, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "tn_free": (C.c_int, [_H, C.c_int]),
    "tn_upload": (C.c_int, [_H, C.c_int, _DP]),
    "tn_download": (C.c_int, [_H, C.c_int, _DP]),
    "tn_contract": (C.c_int, [_H, C.c_double, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_double, C.c_int, C.c_char_p]),
    "tn_add": (C.c_int, [_H, C.c_double, C.c_int, C.c_char_p, C.c_double, C.c_int, C.c_char_p]),
    "tn_dot": (C.c_int, [_H, C.c_int, C.c_int, _DP]),
    "tn_excitation_divide": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]),
    "tn_get_stats": (C.c_int, [_H, _DP, _DP, C.POINTER(C.c_int64)]),
}
_bound = False


class TnError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsisi4s_pt tensor engine error {code}: {msg}")
        self.code = code


def _load():
    global _bound
    lib = _lib.load()
    if not _bound:
        for name, (res, args) in TN_SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _bound = True
    return lib


class Tensor:
    """Handle of one device tensor (id + shape); freed with the engine or by `free()`."""
    __slots__ = ("eng", "id", "shape")

    def __init__(self, eng, id_, shape):
        self.eng, self.id, self.shape = eng, id_, tuple(int(x) for x in shape)

    def get(self) -> np.ndarray:
        return self.eng.download(self)

    def free(self):
        self.eng.free(self)


class DeviceTensors:
    def __init__(self, device: int = 0):
        self.lib = _load()
        self._h = _H()
        self._chk(self.lib.tn_create(C.byref(self._h), int(device)))

    def _chk(self, rc):
        if rc != 0:
            raise TnError(rc, self.lib.tn_last_error().decode())

    def close(self):
        if self._h:
            self.lib.tn_destroy(self._h)
            self._h = _H()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- tensors
    def tensor(self, shape, data=None) -> Tensor:
        shape = tuple(int(x) for x in shape)
        lens = (C.c_int64 * max(1, len(shape)))(*shape)
        tid = C.c_int()
        self._chk(self.lib.tn_tensor(self._h, len(shape), lens, C.byref(tid)))
        t = Tensor(self, tid.value, shape)
        if data is not None:
            self.upload(t, data)
        return t

    def upload(self, t: Tensor, data):
        a = np.asarray(data, dtype=np.float64)
        a = np.asfortranarray(a) if a.ndim else np.ascontiguousarray(a).reshape(())   # asfortranarray makes 0-d arrays 1-d
        if tuple(a.shape) != t.shape:
            raise ValueError(f"expected shape {t.shape}, got {a.shape}")
        self._chk(self.lib.tn_upload(self._h, t.id, a.ctypes.data_as(_DP)))

    def download(self, t: Tensor) -> np.ndarray:
        out = np.zeros(t.shape, dtype=np.float64, order="F")
        self._chk(self.lib.tn_download(self._h, t.id, out.ctypes.data_as(_DP)))
        return out

    def free(self, t: Tensor):
        if t.id >= 0:
            self._chk(self.lib.tn_free(self._h, t.id))
            t.id = -1

    # -- statements
    def contract(self, alpha, A: Tensor, ia: str, B: Tensor, ib: str, beta, Cc: Tensor, ic: str):
        self._chk(self.lib.tn_contract(self._h, float(alpha), A.id, ia.encode(), B.id, ib.encode(), float(beta),
                                       Cc.id, ic.encode()))

    def add(self, alpha, A: Tensor, ia: str, beta, Cc: Tensor, ic: str):
        self._chk(self.lib.tn_add(self._h, float(alpha), A.id, ia.encode(), float(beta), Cc.id, ic.encode()))

    def dot(self, A: Tensor, B: Tensor) -> float:
        out = C.c_double()
        self._chk(self.lib.tn_dot(self._h, A.id, B.id, C.byref(out)))
        return float(out.value)

    def excitation_divide(self, R: Tensor, T: Tensor, epsi: Tensor, epsa: Tensor, shift: float = 0.0):
        self._chk(self.lib.tn_excitation_divide(self._h, R.id, T.id, epsi.id, epsa.id, float(shift)))

    def stats(self):
        f, b, n = C.c_double(), C.c_double(), C.c_int64()
        self._chk(self.lib.tn_get_stats(self._h, C.byref(f), C.byref(b), C.byref(n)))
        return {"flops": f.value, "bytes": b.value, "launches": n.value}
