"""Spin-orbital (unrestricted) perturbative triples on the device -- SURVEY.md 8f, N4.

Reference: UPerturbativeTriples::run (src/algorithms/UPerturbativeTriples.cxx:19-305): the full-tensor
formulation on antisymmetrised integrals, v^3 o^3 intermediates, energy (1/36) DV . T.  Every CTF statement
of the reference is one `contract` / permuted `add` of the device tensor engine below, with the same index
strings; like the reference's version it is meant for small systems (three v^3 o^3 tensors).
"""
from __future__ import annotations

from .tensor_engine import DeviceTensors
from .triples import Algorithm, SisiException, register


def spin_orbital_triples_energy(epsi, epsa, Tai, Tabij, Vabij, Vijka, Vabci, device: int = 0) -> float:
    o, v = int(len(epsi)), int(len(epsa))
    with DeviceTensors(device) as eng:
        ei, ea = eng.tensor((o,), epsi), eng.tensor((v,), epsa)
        t1, t2 = eng.tensor((v, o), Tai), eng.tensor((v, v, o, o), Tabij)
        pphh, hhhp, ppph = eng.tensor((v, v, o, o), Vabij), eng.tensor((o, o, o, v), Vijka), eng.tensor((v, v, v, o), Vabci)
        six = (v, v, v, o, o, o)
        T, DV, SV = eng.tensor(six), eng.tensor(six), eng.tensor(six)
        # VABCI part (:112-124)
        eng.contract(1.0, t2, "adij", ppph, "bcdk", 0.0, DV, "abcijk")
        for sign, idx in ((+1, "defjki"), (-1, "edfjki"), (-1, "fedjki"), (+1, "defkij"), (-1, "edfkij"), (-1, "fedkij"),
                          (+1, "defijk"), (-1, "edfijk"), (-1, "fedijk")):
            eng.add(sign, DV, idx, 1.0, T, "defjki")
        # VIJKA part (:127-137)
        eng.contract(1.0, t2, "deok", hhhp, "ijof", 0.0, DV, "defkij")
        for sign, idx in ((+1, "defkij"), (-1, "dfekij"), (-1, "fedkij"), (-1, "defjik"), (+1, "dfejik"), (+1, "fedjik"),
                          (-1, "defikj"), (+1, "dfeikj"), (+1, "fedikj")):
            eng.add(sign, DV, idx, 1.0, T, "defjki")
        eng.add(1.0, T, "abcijk", 0.0, DV, "abcijk")                      # :140 the antisymmetrised doubles part
        # singles part (:143-153)
        eng.contract(1.0, t1, "dk", pphh, "efij", 0.0, SV, "defkij")
        for sign, idx in ((+1, "defkij"), (-1, "edfkij"), (-1, "fedkij"), (-1, "defjik"), (+1, "edfjik"), (+1, "fedjik"),
                          (-1, "defikj"), (+1, "edfikj"), (+1, "fedikj")):
            eng.add(sign, SV, idx, 1.0, T, "defkij")
        eng.excitation_divide(T, T, ei, ea, 0.0)                           # T / (eps_i+eps_j+eps_k-eps_a-eps_b-eps_c) (:270-283)
        return eng.dot(DV, T) / 36.0                                       # :286


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
