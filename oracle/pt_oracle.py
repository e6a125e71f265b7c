"""CPU oracle for the closed-shell CCSD(T) perturbative-triples energy.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sisi4s_b200/`` may import this file;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
reference legs use it, and only as the checker.

It restates, in NumPy FP64, the algorithm of the reference
(`/root/reference`, alejandrogallo/sisi4s):

  form A  ``triples_loop``     literal ``i<=j<=k`` loop of
          src/algorithms/CcsdPerturbativeTriples.cxx:119-248 (helpers :81-117),
          permutation algebra of src/math/Permutation.hpp:49-101 and the string
          action of CcsdPerturbativeTriples.cxx:22-30.
  form B  ``triples_full``     src/algorithms/PerturbativeTriples.cxx:172-239
          (full v^3 o^3 tensors; defines the PPPHCoulombIntegrals contract).
  form C  ``triples_piecuch``  src/algorithms/PerturbativeTriples.cxx:99-170.

The arithmetic of the reference lives in Cyclops CTF (un-vendored, pinned at
53ae5daad851bf3b198ebe1fa761c13b12291116, configure.ac:40): Einstein-summation
index strings over column-major tensors.  ``A["abc"] += B["bac"]`` means
``A[a,b,c] += B[b,a,c]``; we restate that with ``numpy.einsum`` using the very
same index strings.

PARITY STATUS: PINNED by a known answer the reference repository holds.  The
reference cannot be built here (needs MPI + CTF) and ships no unit test for
this path, but integration-tests/tests/cc4s/ueg/rs1.0-7occ-26virt/
cc4s.correct.out.yaml:124-169 records MP2, CCSD and (T) energies of the
uniform electron gas (rs = 1, 7 occupied / 26 virtual states), a system defined
by closed formulas.  tests/test_known_answers.py regenerates its inputs
(sisi4s_b200/ueg.py, restating UegVertexGenerator.cxx; MP2 agrees to 1e-15),
converges the CCSD amplitudes (oracle/ccsd.py; CCSD energy agrees to 2e-10,
the reference converged to 1e-8) and checks forms A and B and the C port
against the recorded (T) = -0.0063019625641725016: agreement 3e-12.
Additional pins: the permutation tables against the reference's own
``Permutation.hpp`` compiled from /root/reference (oracle/ref_perm_dump.cxx ->
oracle/_ref/), and mutual agreement of the three formulations (A, B, C) the
reference contains on unsymmetric random inputs.  The other known answers of
the reference (H2O, HF molecules) need downloaded input files and are not
reproduced.
"""
from __future__ import annotations

import numpy as np

ORDER = 6


# --------------------------------------------------------------------------
# Permutation<3>, src/math/Permutation.hpp:49-62 (recursive constructor)
# --------------------------------------------------------------------------
def _perm_images(n: int, p: int) -> list[int]:
    """images[] of Permutation<N>(p), restating Permutation.hpp:52-62."""
    if n == 1:
        return [0]  # Permutation<1>, Permutation.hpp:80-84
    sub = _perm_images(n - 1, p // n)
    images = [0] * n
    i = 0
    while i < p % n:
        images[i] = sub[i] + 1
        i += 1
    images[p % n] = 0
    i += 1
    while i < n:
        images[i] = sub[i - 1] + 1
        i += 1
    return images


PERM = [tuple(_perm_images(3, p)) for p in range(ORDER)]
# expected: p0=(0,1,2) p1=(1,0,2) p2=(1,2,0) p3=(0,2,1) p4=(2,0,1) p5=(2,1,0)


def invariant_elements_count(pi) -> int:
    """Permutation.hpp:64-68."""
    return sum(1 for i in range(3) if pi[i] == i)


def map_after(f, tau):
    """(f * tau)(i) = f(tau(i)); Permutation.hpp:96-101 (Map after Permutation)."""
    return tuple(f[tau[i]] for i in range(3))


def str_after(s: str, pi) -> str:
    """std::string * Permutation, CcsdPerturbativeTriples.cxx:22-30."""
    return "".join(s[pi[i]] for i in range(3))


# spinAndFermiFactors, CcsdPerturbativeTriples.cxx:143
SPIN_AND_FERMI = (+2.0, -4.0, 0.0, +8.0)


# --------------------------------------------------------------------------
# form A: literal loop
# --------------------------------------------------------------------------
def doubles_contribution(T2, Vppph, Vhhhp, ijk):
    """getDoublesContribution, CcsdPerturbativeTriples.cxx:87-96, with the
    on-the-fly vertex product replaced by the identical PPPHCoulombIntegrals
    block (CoulombIntegralsFromVertex.cxx:430-431; PerturbativeTriples.cxx:190).
    W[a,b,c] = sum_d T2[a,d,i,j] V[b,c,d,k] - sum_l T2[a,b,i,l] Vhhhp[j,k,l,c]
    """
    i, j, k = ijk
    W = np.einsum("ad,bcd->abc", T2[:, :, i, j], Vppph[:, :, :, k], optimize=True)
    W -= np.einsum("abl,lc->abc", T2[:, :, i, :], Vhhhp[j, k, :, :], optimize=True)
    return W


def singles_contribution(T1, Vpphh, ijk):
    """getSinglesContribution, CcsdPerturbativeTriples.cxx:81-85."""
    i, j, k = ijk
    return 0.5 * np.einsum("a,bc->abc", T1[:, i], Vpphh[:, :, j, k])


def energy_denominator(epsi, epsa, ijk):
    """getEnergyDenominator, CcsdPerturbativeTriples.cxx:98-117."""
    i, j, k = ijk
    D = np.full((epsa.size,) * 3, epsi[i] + epsi[j] + epsi[k])
    D -= epsa[:, None, None]
    D -= epsa[None, :, None]
    D -= epsa[None, None, :]
    return D


def triple_energy(epsi, epsa, T1, T2, Vpphh, Vhhhp, Vppph, ijk):
    """Body of the i<=j<=k loop, CcsdPerturbativeTriples.cxx:159-216, for one
    sorted hole triple.  Returns its contribution to the (T) energy."""
    i = tuple(ijk)
    piDV = [None] * ORDER
    distinct = [False] * ORDER
    DV = np.zeros((epsa.size,) * 3)
    for p in range(ORDER):
        pi = PERM[p]
        q = 0
        while q < p:
            if map_after(i, PERM[q]) == map_after(i, pi):
                break
            q += 1
        if q < p:
            distinct[p] = False
            piDV[p] = piDV[q]
        else:
            distinct[p] = True
            piDV[p] = doubles_contribution(T2, Vppph, Vhhhp, map_after(i, pi))
        # DVabc["abc"] += piDVabc[p]["abc" * pi]
        DV += np.einsum(str_after("abc", pi) + "->abc", piDV[p])
    DV = DV / energy_denominator(epsi, epsa, i)
    e = 0.0
    for p in range(ORDER):
        if not distinct[p]:
            continue
        pi = PERM[p]
        Tabc = np.zeros_like(DV)
        SV = singles_contribution(T1, Vpphh, map_after(i, pi))
        for s in range(ORDER):
            sigma = PERM[s]
            sf = SPIN_AND_FERMI[invariant_elements_count(sigma)]
            idx = str_after(str_after("abc", sigma), pi)  # ("abc"*sigma)*pi
            Tabc += sf * np.einsum(idx + "->abc", piDV[p])
            Tabc += sf * np.einsum(idx + "->abc", SV)
        e += float(np.einsum("abc,abc->", DV, Tabc))
    return e


def sorted_triples(o: int):
    """Enumeration order of CcsdPerturbativeTriples.cxx:156-158."""
    return [(i, j, k) for i in range(o) for j in range(i, o) for k in range(j, o)]


def triples_loop(epsi, epsa, T1, T2, Vpphh, Vhhhp, Vppph, return_per_triple=False):
    """Form A.  E(T) = sum over sorted triples (CcsdPerturbativeTriples.cxx:240)."""
    o = epsi.size
    es = []
    for ijk in sorted_triples(o):
        es.append(triple_energy(epsi, epsa, T1, T2, Vpphh, Vhhhp, Vppph, ijk))
    total = float(np.sum(np.array(es, dtype=np.longdouble)))
    if return_per_triple:
        return total, np.array(es)
    return total


# --------------------------------------------------------------------------
# form B: full tensors, PerturbativeTriples.cxx:172-239
# --------------------------------------------------------------------------
def _denominator6(epsi, epsa):
    o, v = epsi.size, epsa.size
    D = np.zeros((v, v, v, o, o, o))
    D += epsi[None, None, None, :, None, None]
    D += epsi[None, None, None, None, :, None]
    D += epsi[None, None, None, None, None, :]
    D -= epsa[:, None, None, None, None, None]
    D -= epsa[None, :, None, None, None, None]
    D -= epsa[None, None, :, None, None, None]
    return D


def triples_full(epsi, epsa, T1, T2, Vpphh, Vhhhp, Vppph):
    SV = 0.5 * np.einsum("ai,bcjk->abcijk", T1, Vpphh)
    DV = np.einsum("bcdk,adij->abcijk", Vppph, T2, optimize=True)
    DV -= np.einsum("jklc,abil->abcijk", Vhhhp, T2, optimize=True)
    T = np.zeros_like(DV)
    for f, s in ((8.0, "abcijk"), (-4.0, "acbijk"), (-4.0, "bacijk"),
                 (2.0, "bcaijk"), (2.0, "cabijk"), (-4.0, "cbaijk")):
        T += f * np.einsum(s + "->abcijk", DV)
        T += f * np.einsum(s + "->abcijk", SV)
    T = T / _denominator6(epsi, epsa)
    e = 0.0
    for s in ("abcijk", "bacjik", "acbikj", "cbakji", "cabkij", "bcajki"):
        e += float(np.einsum(s + ",abcijk->", DV, T))
    return e


# --------------------------------------------------------------------------
# form C: Piecuch factors, PerturbativeTriples.cxx:99-170
# --------------------------------------------------------------------------
def triples_piecuch(epsi, epsa, T1, T2, Vpphh, Vhhhp, Vppph):
    T = np.einsum("bcek,aeij->abcijk", Vppph, T2, optimize=True)
    T -= np.einsum("jkmc,abim->abcijk", Vhhhp, T2, optimize=True)
    X = T.copy()
    for s in ("bacjik", "acbikj", "cbakji", "cabkij", "bcajki"):
        T += np.einsum(s + "->abcijk", X)
    Z = np.einsum("ai,bcjk->abcijk", T1, Vpphh)
    Z += np.einsum("bj,acik->abcijk", T1, Vpphh)
    Z += np.einsum("ck,abij->abcijk", T1, Vpphh)
    X = (4.0 / 3.0) * Z
    X += (-2.0) * np.einsum("acbijk->abcijk", Z)
    X += (2.0 / 3.0) * np.einsum("bcaijk->abcijk", Z)
    X += (4.0 / 3.0) * T
    X += (-2.0) * np.einsum("acbijk->abcijk", T)
    X += (2.0 / 3.0) * np.einsum("bcaijk->abcijk", T)
    T = T / _denominator6(epsi, epsa)
    return float(np.einsum("abcijk,abcijk->", X, T))


# --------------------------------------------------------------------------
# integrals from a Coulomb vertex, CoulombIntegralsFromVertex.cxx:121-136,399-433
# --------------------------------------------------------------------------
def integrals_from_vertex(Gamma, o, v):
    """Gamma[F,p,q] complex, Np=o+v states, holes first, particles = last v
    (CoulombIntegralsFromVertex.cxx:121-136).  Real-integral formulas:
      Vabij["abij"] = Re G["Gai"] Re G["Gbj"] + Im Im          (:402-403)
      Vijka["ijka"] = Re G["Gik"] Re G["Gaj"] + Im Im          (:416-417)
      Vabci["abci"] = Re G["Gac"] Re G["Gbi"] + Im Im          (:430-431)
    """
    Np = Gamma.shape[1]
    a0 = Np - v
    Gij = Gamma[:, :o, :o]
    Gai = Gamma[:, a0:, :o]
    Gab = Gamma[:, a0:, a0:]

    def rr(x, sx, y, sy, out):
        return (np.einsum(f"{sx},{sy}->{out}", x.real, y.real, optimize=True)
                + np.einsum(f"{sx},{sy}->{out}", x.imag, y.imag, optimize=True))

    Vpphh = rr(Gai, "Gai", Gai, "Gbj", "abij")
    Vhhhp = rr(Gij, "Gik", Gai, "Gaj", "ijka")
    Vppph = rr(Gab, "Gac", Gai, "Gbi", "abci")
    return Vpphh, Vhhhp, Vppph
