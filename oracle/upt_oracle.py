"""NumPy restatement of the reference's spin-orbital (unrestricted) perturbative triples -- TEST INFRASTRUCTURE.

Literal restatement (same index strings, same order) of UPerturbativeTriples::run (reference
src/algorithms/UPerturbativeTriples.cxx:19-305): full v^3 o^3 tensors, antisymmetrised integrals
`<ab||ij>`, `<ij||ka>`, `<ab||ci>` and amplitudes as inputs, energy (1/36) DV . T.  PARITY: the reference
holds no known answer for this step (it is not even compiled, src/Makefile.am:130); the restatement is
pinned by reproducing the CLOSED-SHELL (T) energy (oracle/pt_oracle.py, itself pinned by the recorded UEG
answer) when it is fed the spin-orbital image of a closed-shell system (tests/test_upt.py).
Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np


def es(spec, *ops):
    return np.einsum(spec, *ops, optimize=True)


def triples(epsi, epsa, Tai, Tabij, Vabij, Vijka, Vabci):
    """UPerturbativeTriples::run (:19-305); returns the triples energy (PerturbativeTriplesEnergy, :305)."""
    T = np.zeros((len(epsa),) * 3 + (len(epsi),) * 3)
    # VABCI part (:112-124): Tabcijk["defjki"] accumulates permutations of DVabcijk
    DV = es("adij,bcdk->abcijk", Tabij, Vabci)
    for sign, idx in ((+1, "defjki"), (-1, "edfjki"), (-1, "fedjki"), (+1, "defkij"), (-1, "edfkij"), (-1, "fedkij"),
                      (+1, "defijk"), (-1, "edfijk"), (-1, "fedijk")):
        T += sign * es(f"{idx}->defjki", DV)
    # VIJKA part (:127-137): DVabcijk["defkij"] = Tabij["deok"] Vijka["ijof"]
    DV = es("deok,ijof->defkij", Tabij, Vijka)
    for sign, idx in ((+1, "defkij"), (-1, "dfekij"), (-1, "fedkij"), (-1, "defjik"), (+1, "dfejik"), (+1, "fedjik"),
                      (-1, "defikj"), (+1, "dfeikj"), (+1, "fedikj")):
        T += sign * es(f"{idx}->defjki", DV)
    DVs = T.copy()                                                   # :140  DVabcijk["abcijk"] = Tabcijk["abcijk"]
    # singles part (:143-153): SVabcijk["defkij"] = Tai["dk"] Vabij["efij"]
    SV = es("dk,efij->defkij", Tai, Vabij)
    for sign, idx in ((+1, "defkij"), (-1, "edfkij"), (-1, "fedkij"), (-1, "defjik"), (+1, "edfjik"), (+1, "fedjik"),
                      (-1, "defikj"), (+1, "edfikj"), (+1, "fedikj")):
        T += sign * es(f"{idx}->defkij", SV)
    D = (epsi[None, None, None, :, None, None] + epsi[None, None, None, None, :, None] + epsi[None, None, None, None, None, :]
         - epsa[:, None, None, None, None, None] - epsa[None, :, None, None, None, None] - epsa[None, None, :, None, None, None])
    T = T / D                                                        # :270-283
    return float((1.0 / 36.0) * np.sum(DVs * T))                     # :286


def spin_orbital_image(epsi, epsa, T1, T2, gamma):
    """Spin-orbital (alpha block, then beta block) antisymmetrised image of a closed-shell system:
    eigenenergies, T1[A,I], T2[A,B,I,J], <AB||IJ>, <IJ||KA>, <AB||CI> with <pq|rs> = G[p,r].G[q,s]."""
    o, v = len(epsi), len(epsa)
    np_ = gamma.shape[1]
    V = es("Gpr,Gqs->pqrs", gamma.real, gamma.real) + es("Gpr,Gqs->pqrs", gamma.imag, gamma.imag)
    h, p = np.arange(o), np.arange(np_ - v, np_)
    occ = [(x, s) for s in (0, 1) for x in h]
    vir = [(x, s) for s in (0, 1) for x in p]

    def anti(P, Q, R, S):
        out = np.zeros((len(P), len(Q), len(R), len(S)))
        for a, (pa, sa) in enumerate(P):
            for b, (pb, sb) in enumerate(Q):
                for c, (pc, sc) in enumerate(R):
                    for d, (pd, sd) in enumerate(S):
                        x = V[pa, pb, pc, pd] if (sa == sc and sb == sd) else 0.0
                        y = V[pa, pb, pd, pc] if (sa == sd and sb == sc) else 0.0
                        out[a, b, c, d] = x - y
        return out

    O, Vv = 2 * o, 2 * v
    T1u = np.zeros((Vv, O))
    T2u = np.zeros((Vv, Vv, O, O))
    for A, (a, sa) in enumerate(vir):
        for I, (i, si) in enumerate(occ):
            if sa == si:
                T1u[A, I] = T1[a - (np_ - v), i]
            for B, (b, sb) in enumerate(vir):
                for J, (j, sj) in enumerate(occ):
                    x = T2[a - (np_ - v), b - (np_ - v), i, j] if (sa == si and sb == sj) else 0.0
                    y = T2[b - (np_ - v), a - (np_ - v), i, j] if (sb == si and sa == sj) else 0.0
                    T2u[A, B, I, J] = x - y
    return (np.tile(epsi, 2), np.tile(epsa, 2), T1u, T2u, anti(vir, vir, occ, occ), anti(occ, occ, occ, vir),
            anti(vir, vir, vir, occ))
