"""NumPy restatement of the reference's closed-shell CCSD solver -- TEST INFRASTRUCTURE (oracle side).

Line-by-line (same index strings, same order of terms) restatement of

  * CcsdEnergyFromCoulombIntegralsReference::getResiduum   (reference
    src/algorithms/CcsdEnergyFromCoulombIntegralsReference.cxx:29-295, Hirata et al. CPL 345, 475 (2001)),
  * ClusterSinglesDoublesAlgorithm::run / getEnergy / estimateAmplitudesFromResiduum /
    calculateExcitationEnergies (src/algorithms/ClusterSinglesDoublesAlgorithm.cxx:37-128, 130-205,
    302-331, 343-365),
  * LinearMixer (src/mixers/LinearMixer.cxx:31-49) and DiisMixer (src/mixers/DiisMixer.cxx:103-181,
    the small symmetric solve :16-41 done with numpy.linalg.solve instead of dsysv_).

CTF semantics: `A["abij"] += B["cdkl"] * C["adkl"]` sums every index that does not appear on the left.
Everything is written as numpy.einsum on the same strings.  PARITY: pinned by the CCSD energy the
reference records for the UEG test system (integration-tests/tests/cc4s/ueg/rs1.0-7occ-26virt/
cc4s.correct.out.yaml:153, -0.39269658954585018) in tests/test_ccsd_ref.py.

Only tests/ may import this module; the product solver is sisi4s_b200/ccsd.py (device kernels).
"""
from __future__ import annotations

import numpy as np


def es(spec, *ops):
    return np.einsum(spec, *ops, optimize=True)


def residuum(i, Tai, Tabij, V, initial_doubles_given=False):
    """getResiduum(i, amplitudes) (:29-295).  V: dict with PPHH, PHPH, HHHH, HHHP, PPPH, PPPP blocks."""
    Vabij = V["PPHH"]
    Rai = np.zeros_like(Tai)
    if i == 0 and not initial_doubles_given:
        return Rai, Vabij.copy()                                                    # :52-57 MP2 amplitudes
    Vabcd, Vaibj, Vijkl, Vijka, Vabci = V["PPPP"], V["PHPH"], V["HHHH"], V["HHHP"], V["PPPH"]
    Rabij = np.zeros_like(Tabij)
    # Kac (:169-173)
    Kac = -2.0 * es("cdkl,adkl->ac", Vabij, Tabij)
    Kac += 1.0 * es("dckl,adkl->ac", Vabij, Tabij)
    Kac += -2.0 * es("cdkl,ak,dl->ac", Vabij, Tai, Tai)
    Kac += 1.0 * es("dckl,ak,dl->ac", Vabij, Tai, Tai)
    # Lac (:176-178)
    Lac = Kac.copy()
    Lac += 2.0 * es("cdak,dk->ac", Vabci, Tai)
    Lac += -1.0 * es("dcak,dk->ac", Vabci, Tai)
    # Kki (:181-184)
    Kki = 2.0 * es("cdkl,cdil->ki", Vabij, Tabij)
    Kki += -1.0 * es("dckl,cdil->ki", Vabij, Tabij)
    Kki += 2.0 * es("cdkl,ci,dl->ki", Vabij, Tai, Tai)
    Kki += -1.0 * es("dckl,ci,dl->ki", Vabij, Tai, Tai)
    # Lki (:187-189)
    Lki = Kki.copy()
    Lki += 2.0 * es("klic,cl->ki", Vijka, Tai)
    Lki += -1.0 * es("lkic,cl->ki", Vijka, Tai)
    # :192-201
    Rabij += 1.0 * es("ac,cbij->abij", Lac, Tabij)
    Rabij += -1.0 * es("ki,abkj->abij", Lki, Tabij)
    Rabij += 1.0 * es("baci,cj->abij", Vabci, Tai)
    Rabij += -1.0 * es("bkci,ak,cj->abij", Vaibj, Tai, Tai)
    Rabij += -1.0 * es("jika,bk->abij", Vijka, Tai)
    Rabij += -1.0 * es("acik,cj,bk->abij", Vabij, Tai, Tai)
    # Xakic (:204-210)
    Xakic = es("acik->akic", Vabij).copy()
    Xakic += -1.0 * es("lkic,al->akic", Vijka, Tai)
    Xakic += 1.0 * es("acdk,di->akic", Vabci, Tai)
    Xakic += -0.5 * es("dclk,dail->akic", Vabij, Tabij)
    Xakic += -1.0 * es("dclk,di,al->akic", Vabij, Tai, Tai)
    Xakic += 1.0 * es("dclk,adil->akic", Vabij, Tabij)
    Xakic += -0.5 * es("cdlk,adil->akic", Vabij, Tabij)
    # Xakci (:213-217)
    Xakci = Vaibj.copy()
    Xakci += -1.0 * es("klic,al->akci", Vijka, Tai)
    Xakci += 1.0 * es("adck,di->akci", Vabci, Tai)
    Xakci += -0.5 * es("cdlk,dail->akci", Vabij, Tabij)
    Xakci += -1.0 * es("cdlk,di,al->akci", Vabij, Tai, Tai)
    # :220-224
    Rabij += 2.0 * es("akic,cbkj->abij", Xakic, Tabij)
    Rabij += -1.0 * es("akic,bckj->abij", Xakic, Tabij)
    Rabij += -1.0 * es("akci,cbkj->abij", Xakci, Tabij)
    Rabij += -1.0 * es("bkci,ackj->abij", Xakci, Tabij)
    # symmetrise with the permutation operator (:228-229)
    Rabij = Rabij + es("abij->baji", Rabij)
    # :238
    Rabij += Vabij
    # Xklij (:241-245)
    Xklij = Vijkl.copy()
    Xklij += es("klic,cj->klij", Vijka, Tai)
    Xklij += es("lkjc,ci->klij", Vijka, Tai)
    Xklij += es("cdkl,cdij->klij", Vabij, Tabij)
    Xklij += es("cdkl,ci,dj->klij", Vabij, Tai, Tai)
    # :248-251
    Rabij += es("klij,abkl->abij", Xklij, Tabij)
    Rabij += es("klij,ak,bl->abij", Xklij, Tai, Tai)
    # Xabcd (:254-256)
    Xabcd = 1.0 * Vabcd
    Xabcd = Xabcd + -1.0 * es("cdak,bk->abcd", Vabci, Tai)
    Xabcd += -1.0 * es("dcbk,ak->abcd", Vabci, Tai)
    # :259-260
    Rabij += es("abcd,cdij->abij", Xabcd, Tabij)
    Rabij += es("abcd,ci,dj->abij", Xabcd, Tai, Tai)
    # T1 equations (:270-293)
    Rai += 1.0 * es("ac,ci->ai", Kac, Tai)
    Rai += -1.0 * es("ki,ak->ai", Kki, Tai)
    Kck = 2.0 * es("cdkl,dl->ck", Vabij, Tai)
    Kck += -1.0 * es("cdlk,dl->ck", Vabij, Tai)
    Rai += 2.0 * es("ck,caki->ai", Kck, Tabij)
    Rai += -1.0 * es("ck,caik->ai", Kck, Tabij)
    Rai += 1.0 * es("ck,ci,ak->ai", Kck, Tai, Tai)
    Rai += 2.0 * es("acik,ck->ai", Vabij, Tai)
    Rai += -1.0 * es("akci,ck->ai", Vaibj, Tai)
    Rai += 2.0 * es("cdak,cdik->ai", Vabci, Tabij)
    Rai += -1.0 * es("dcak,cdik->ai", Vabci, Tabij)
    Rai += 2.0 * es("cdak,ci,dk->ai", Vabci, Tai, Tai)
    Rai += -1.0 * es("dcak,ci,dk->ai", Vabci, Tai, Tai)
    Rai += -2.0 * es("klic,ackl->ai", Vijka, Tabij)
    Rai += 1.0 * es("lkic,ackl->ai", Vijka, Tabij)
    Rai += -2.0 * es("klic,ak,cl->ai", Vijka, Tai, Tai)
    Rai += 1.0 * es("lkic,ak,cl->ai", Vijka, Tai, Tai)
    return Rai, Rabij


def energy(Tai, Tabij, Vabij):
    """getEnergy (:130-205), closed shell (spins = 2), not antisymmetrised: direct + exchange."""
    dire = 0.5 * 4.0 * (es("abij,abij->", Tabij, Vabij) + es("ai,bj,abij->", Tai, Tai, Vabij))
    exce = -0.5 * 2.0 * (es("abij,baij->", Tabij, Vabij) + es("ai,bj,baij->", Tai, Tai, Vabij))
    return float(dire + exce)


def estimate_amplitudes(Rai, Rabij, Tai, Tabij, epsi, epsa, level_shift=0.0):
    """estimateAmplitudesFromResiduum (:302-331): R -= shift * T; R = -R / (D + shift),
    D = sum eps_a - sum eps_i (calculateExcitationEnergies :343-365)."""
    D1 = epsa[:, None] - epsi[None, :]
    D2 = (epsa[:, None, None, None] + epsa[None, :, None, None]
          - epsi[None, None, :, None] - epsi[None, None, None, :])
    return (-(Rai - level_shift * Tai) / (D1 + level_shift),
            -(Rabij - level_shift * Tabij) / (D2 + level_shift))


class LinearMixer:
    def __init__(self, ratio=1.0):
        self.ratio, self.last = ratio, None

    def append(self, A, R):
        if self.last is not None:
            A = [self.ratio * a + (1 - self.ratio) * l for a, l in zip(A, self.last)]
        self.last = A

    def get(self):
        return self.last


class DiisMixer:
    """DiisMixer.cxx:55-181: B matrix with the -1 border, overlaps 2 Re <R_i|R_j>, first column of B^-1."""
    def __init__(self, max_residua=4):
        N = self.N = int(max_residua)
        self.amplitudes, self.residua = [None] * N, [None] * N
        self.next_index = self.count = 0
        self.B = np.zeros((N + 1, N + 1))
        self.B[0, 1:] = -1.0
        self.B[1:, 0] = -1.0
        self.next = None

    def append(self, A, R):
        N, n = self.N, self.next_index
        self.amplitudes[n], self.residua[n] = A, R
        for i in range(N):
            if self.residua[i] is not None:
                ov = 2.0 * sum(float(np.vdot(x, y)) for x, y in zip(self.residua[i], R))
                self.B[n + 1, i + 1] = self.B[i + 1, n + 1] = ov
        if self.count < N:
            self.count += 1
        dim = self.count + 1
        rhs = np.zeros(dim)
        rhs[0] = -1.0
        col = np.linalg.solve(self.B[:dim, :dim], rhs)
        self.next = [np.zeros_like(a) for a in A]
        for j in range(self.count):
            i = (n + N - j) % N
            for t, a in zip(self.next, self.amplitudes[i]):
                t += col[i + 1] * a
        self.next_index = (n + 1) % N

    def get(self):
        return self.next


def solve(epsi, epsa, V, mixer="LinearMixer", max_residua=4, mixing_ratio=1.0, max_iterations=16,
          energy_convergence=1e-6, amplitudes_convergence=1e-5, level_shift=0.0, log=None):
    """ClusterSinglesDoublesAlgorithm::run<double> (:37-128); defaults as in the reference header.
    Returns dict(energy, T1, T2, iterations, converged)."""
    nv, no = len(epsa), len(epsi)
    Tai, Tabij = np.zeros((nv, no)), np.zeros((nv, nv, no, no))
    mix = DiisMixer(max_residua) if mixer == "DiisMixer" else LinearMixer(mixing_ratio)
    e = prev = 0.0
    it, converged = 0, False
    for it in range(max_iterations):
        Rai, Rabij = residuum(it, Tai, Tabij, V)
        Eai, Eabij = estimate_amplitudes(Rai, Rabij, Tai, Tabij, epsi, epsa, level_shift)
        dai, dabij = Eai - Tai, Eabij - Tabij
        mix.append([Eai, Eabij], [dai, dabij])
        Tai, Tabij = mix.get()
        e = energy(Tai, Tabij, V["PPHH"])
        if log:
            log(f"iteration {it + 1}: energy = {e:.12f}")
        dd = float(np.vdot(dai, dai) + np.vdot(dabij, dabij))
        tt = float(np.vdot(Tai, Tai) + np.vdot(Tabij, Tabij))
        if abs((e - prev) / e) < abs(energy_convergence) and abs(dd / tt) < abs(amplitudes_convergence ** 2):
            converged = True
            break
        prev = e
    return dict(energy=e, T1=Tai, T2=Tabij, iterations=it + 1, converged=converged)


def integral_blocks(gamma, no, nv):
    """The six blocks getResiduum reads, with the reference's index strings
    (CoulombIntegralsFromVertex.cxx:395-431)."""
    np_ = gamma.shape[1]
    h, p = slice(0, no), slice(np_ - nv, np_)
    parts = {"ij": gamma[:, h, h], "ai": gamma[:, p, h], "ab": gamma[:, p, p]}

    def block(a, ia, b, ib, out):
        return (es(f"G{ia},G{ib}->{out}", parts[a].real, parts[b].real)
                + es(f"G{ia},G{ib}->{out}", parts[a].imag, parts[b].imag))
    return {"PPHH": block("ai", "ai", "ai", "bj", "abij"), "HHHH": block("ij", "ik", "ij", "jl", "ijkl"),
            "HHHP": block("ij", "ik", "ai", "aj", "ijka"), "PPPP": block("ab", "ac", "ab", "bd", "abcd"),
            "PPPH": block("ab", "ac", "ai", "bi", "abci"), "PHPH": block("ab", "ab", "ij", "ij", "aibj")}
