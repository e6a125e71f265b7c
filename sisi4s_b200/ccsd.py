"""ctypes binding of the device closed-shell CCSD solver (include/sisi4s_ccsd.h, csrc/ccsd_solver.cu) -- the
step in front of the (T) path.

The solver itself lives in the library: the reference's residuum
(CcsdEnergyFromCoulombIntegralsReference.cxx:29-295) statement by statement on the device tensor engine,
the solver loop / convergence test / amplitude update of ClusterSinglesDoublesAlgorithm.cxx:37-128,302-331
and the Linear / DIIS mixers.  This module only moves arrays across the C ABI; the plan step
`CcsdEnergyFromCoulombIntegrals[Reference]` (plan.py) and the C++ plugin class
(csrc/CcsdEnergyFromCoulombIntegralsGpu.cxx) are its two callers.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .tensor_engine import TnError, _load as _load_tn

DEFAULT_MAX_ITERATIONS = 16          # ClusterSinglesDoublesAlgorithm.hpp
DEFAULT_ENERGY_CONVERGENCE = 1e-6
DEFAULT_AMPLITUDES_CONVERGENCE = 1e-5
DEFAULT_LEVEL_SHIFT = 0.0
BLOCKS = ("PPHH", "PHPH", "HHHH", "HHHP", "PPPH", "PPPP")
MIXERS = {"LinearMixer": 0, "DiisMixer": 1}

_DP = C.POINTER(C.c_double)
_H = C.c_void_p


class CcsdOptions(C.Structure):
    _fields_ = [("mixer", C.c_int32), ("max_residua", C.c_int32), ("mixing_ratio", C.c_double),
                ("max_iterations", C.c_int32), ("reserved", C.c_int32), ("energy_convergence", C.c_double),
                ("amplitudes_convergence", C.c_double), ("level_shift", C.c_double)]


class CcsdResult(C.Structure):
    _fields_ = [("energy", C.c_double), ("direct", C.c_double), ("exchange", C.c_double), ("iterations", C.c_int32),
                ("converged", C.c_int32), ("flops", C.c_double), ("kernel_launches", C.c_int64)]


CCSD_SYMBOLS = {
    "ccsd_create": (C.c_int, [C.POINTER(_H), C.c_int, C.c_int, C.c_int]),
    "ccsd_destroy": (C.c_int, [_H]),
    "ccsd_default_options": (None, [C.POINTER(CcsdOptions)]),
    "ccsd_set_eigenenergies": (C.c_int, [_H, _DP, _DP]),
    "ccsd_set_integrals": (C.c_int, [_H, C.c_char_p, _DP]),
    "ccsd_set_vertex": (C.c_int, [_H, C.c_int, C.c_int, _DP, _DP]),
    "ccsd_get_integrals": (C.c_int, [_H, C.c_char_p, _DP]),
    "ccsd_set_amplitudes": (C.c_int, [_H, _DP, _DP]),
    "ccsd_residuum": (C.c_int, [_H, C.c_int, _DP, _DP]),
    "ccsd_solve": (C.c_int, [_H, C.POINTER(CcsdOptions), C.POINTER(CcsdResult)]),
    "ccsd_get_amplitudes": (C.c_int, [_H, _DP, _DP]),
}
_bound = False


def _load():
    global _bound
    lib = _load_tn()
    if not _bound:
        for name, (res, args) in CCSD_SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _bound = True
    return lib


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_DP)


class CcsdSolver:
    """One `ccsd_handle_t`: device-resident integrals + amplitudes of a closed-shell CCSD calculation."""

    def __init__(self, epsi, epsa, integrals: dict | None = None, device: int = 0, vertex=None):
        """integrals: the blocks getResiduum reads (:49,136-140) as column-major arrays PPHH[v,v,o,o],
        PHPH[v,o,v,o], HHHH[o,o,o,o], HHHP[o,o,o,v], PPPH[v,v,v,o], PPPP[v,v,v,v]; or `vertex`
        (CoulombVertex[NF,Np,Np] complex): the blocks are then built on the device."""
        self.lib = _load()
        self.no, self.nv = int(len(epsi)), int(len(epsa))
        self._h = _H()
        self._chk(self.lib.ccsd_create(C.byref(self._h), self.no, self.nv, int(device)))
        ei, ea = _f(epsi), _f(epsa)
        self._chk(self.lib.ccsd_set_eigenenergies(self._h, _p(ei), _p(ea)))
        if vertex is not None:
            g = np.asarray(vertex)
            gre, gim = _f(g.real), _f(g.imag)
            self._chk(self.lib.ccsd_set_vertex(self._h, g.shape[0], g.shape[1], _p(gre), _p(gim)))
        for name, block in (integrals or {}).items():
            if name in BLOCKS:
                self.set_integrals(name, block)
        if vertex is None:
            missing = [b for b in BLOCKS if b not in (integrals or {})]
            if missing:
                self.close()
                raise ValueError("Missing argument: " + ", ".join(m + "CoulombIntegrals" for m in missing))

    def _chk(self, rc):
        if rc != 0:
            raise TnError(rc, self.lib.tn_last_error().decode())

    def close(self):
        if self._h:
            self.lib.ccsd_destroy(self._h)
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

    def _shape(self, name):
        return tuple(self.nv if c == "P" else self.no for c in name)

    def set_integrals(self, name: str, block):
        b = _f(block)
        if tuple(b.shape) != self._shape(name):
            raise ValueError(f"{name}CoulombIntegrals: expected shape {self._shape(name)}, got {tuple(b.shape)}")
        self._chk(self.lib.ccsd_set_integrals(self._h, name.encode(), _p(b)))

    def get_integrals(self, name: str) -> np.ndarray:
        out = np.zeros(self._shape(name), order="F")
        self._chk(self.lib.ccsd_get_integrals(self._h, name.encode(), _p(out)))
        return out

    def set_amplitudes(self, T1=None, T2=None):
        """initialSinglesAmplitudes / initialDoublesAmplitudes (createAmplitudes :207-237)."""
        t1 = _f(T1) if T1 is not None else None
        t2 = _f(T2) if T2 is not None else None
        self._chk(self.lib.ccsd_set_amplitudes(self._h, _p(t1) if t1 is not None else None,
                                                _p(t2) if t2 is not None else None))

    def amplitudes(self):
        t1 = np.zeros((self.nv, self.no), order="F")
        t2 = np.zeros((self.nv, self.nv, self.no, self.no), order="F")
        self._chk(self.lib.ccsd_get_amplitudes(self._h, _p(t1), _p(t2)))
        return t1, t2

    def residuum(self, iteration: int):
        """getResiduum(iteration, current amplitudes) -> (Rai, Rabij)."""
        r1 = np.zeros((self.nv, self.no), order="F")
        r2 = np.zeros((self.nv, self.nv, self.no, self.no), order="F")
        self._chk(self.lib.ccsd_residuum(self._h, int(iteration), _p(r1), _p(r2)))
        return r1, r2

    def solve(self, mixer="LinearMixer", max_residua=4, mixing_ratio=1.0, max_iterations=DEFAULT_MAX_ITERATIONS,
              energy_convergence=DEFAULT_ENERGY_CONVERGENCE, amplitudes_convergence=DEFAULT_AMPLITUDES_CONVERGENCE,
              level_shift=DEFAULT_LEVEL_SHIFT):
        """ClusterSinglesDoublesAlgorithm::run<double> (:37-128).  Returns dict(energy, direct, exchange, T1, T2,
        iterations, converged, stats); not converging is reported, not raised (the reference logs a WARNING
        and stores the amplitudes, :120-124)."""
        if mixer not in MIXERS:
            raise ValueError(f"Mixer not implemented: {mixer}")                 # :50-54
        opt = CcsdOptions(MIXERS[mixer], int(max_residua), float(mixing_ratio), int(max_iterations), 0,
                          float(energy_convergence), float(amplitudes_convergence), float(level_shift))
        res = CcsdResult()
        self._chk(self.lib.ccsd_solve(self._h, C.byref(opt), C.byref(res)))
        t1, t2 = self.amplitudes()
        return dict(energy=float(res.energy), direct=float(res.direct), exchange=float(res.exchange), T1=t1, T2=t2,
                    iterations=int(res.iterations), converged=bool(res.converged),
                    stats={"flops": float(res.flops), "launches": int(res.kernel_launches)})


def solve_ccsd(epsi, epsa, integrals=None, device: int = 0, vertex=None, **kw):
    """One-call form: integral blocks or the vertex (host arrays) -> converged amplitudes (host arrays)."""
    with CcsdSolver(epsi, epsa, integrals, device, vertex=vertex) as s:
        return s.solve(**kw)
