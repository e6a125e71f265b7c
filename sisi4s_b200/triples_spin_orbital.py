"""Spin-orbital (unrestricted) perturbative triples on the device -- SURVEY.md 8f, N4.

Reference: UPerturbativeTriples::run (src/algorithms/UPerturbativeTriples.cxx:19-305): the full-tensor
formulation on antisymmetrised integrals, v^3 o^3 intermediates, energy (1/36) DV . T.  Every CTF statement
of the reference is one `tn_contract` / permuted `tn_add` of the device tensor engine (csrc/upt.cu), with the same
index strings; like the reference's version it is meant for small systems (three v^3 o^3 tensors).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .triples import Algorithm, SisiException, register


def spin_orbital_triples_energy(epsi, epsa, Tai, Tabij, Vabij, Vijka, Vabci, device: int = 0) -> float:
    """Triples energy of the spin-orbital full-tensor form through the C ABI (pt_spin_orbital_triples,
    csrc/upt.cu: the reference's statements on the device tensor engine)."""
    o, v = int(len(epsi)), int(len(epsa))
    f = lambda a: np.asfortranarray(a, dtype=np.float64)
    arrs = []
    for a, shape, name in ((epsi, (o,), "HoleEigenEnergies"), (epsa, (v,), "ParticleEigenEnergies"),
                           (Tai, (v, o), "CcsdSinglesAmplitudes"), (Tabij, (v, v, o, o), "CcsdDoublesAmplitudes"),
                           (Vabij, (v, v, o, o), "PPHHCoulombIntegrals"), (Vijka, (o, o, o, v), "HHHPCoulombIntegrals"),
                           (Vabci, (v, v, v, o), "PPPHCoulombIntegrals")):
        a = f(a)
        if tuple(a.shape) != shape:
            raise ValueError(f"{name}: expected shape {shape}, got {tuple(a.shape)}")
        arrs.append(a)
    e = C.c_double(0.0)
    _lib.check(_lib.load().pt_spin_orbital_triples(o, v, int(device), *[a.ctypes.data_as(C.POINTER(C.c_double)) for a in arrs],
                                                   C.byref(e)))
    return float(e.value)


@register
class UPerturbativeTriples(Algorithm):
    """Plan step under the reference's name and keys (UPerturbativeTriples.cxx:19-27,305): the output
    PerturbativeTriplesEnergy is the triples energy alone (the reference has the CcsdEnergy sum commented out)."""
    name = "UPerturbativeTriples"

    def run(self):
        e = spin_orbital_triples_energy(
            self.getTensorArgument("HoleEigenEnergies"), self.getTensorArgument("ParticleEigenEnergies"),
            self.getTensorArgument("CcsdSinglesAmplitudes"), self.getTensorArgument("CcsdDoublesAmplitudes"),
            self.getTensorArgument("PPHHCoulombIntegrals"), self.getTensorArgument("HHHPCoulombIntegrals"),
            self.getTensorArgument("PPPHCoulombIntegrals"), device=self.getIntegerArgument("device", 0))
        self.log = {"triples": e}
        if not self.isArgumentGiven("PerturbativeTriplesEnergy"):
            raise SisiException("Missing argument: PerturbativeTriplesEnergy")
        self.setRealArgument("PerturbativeTriplesEnergy", e)
        return e
