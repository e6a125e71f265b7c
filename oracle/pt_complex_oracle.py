"""NumPy restatement of the reference's COMPLEX closed-shell perturbative triples -- TEST INFRASTRUCTURE.

Follows CcsdPerturbativeTriplesComplex::Calculator<complex>::calculate (reference
src/algorithms/CcsdPerturbativeTriplesComplex.cxx:166-271) with

  * getDoublesParticleContribution<complex> (:341-348):
        DVabc["abc"]  = Tabij(i0,i1)["adij"] * conj(GammaFab)["Fdb"] * GammaFai(i2)["Fck"]
  * addDoublesHoleContribution (:135-140):
        DVabc["abc"] -= Tabil(i0)["abil"] * Valij(i2,i1)["clkj"]          (PHHHCoulombIntegrals)
  * getSinglesContribution (:142-148):  SVabc["abc"] = 0.5 Tai(i0)["ai"] Vabij(i1,i2)["bcjk"]
  * the division DVabc <- conj(DVabc / Delta) (:224-230), the spin factors {+2,-4,0,+8}[invariant
    elements] (:187,241) and the permutation algebra of src/math/Permutation.hpp:49-101.

The energy is the COMPLEX sum of DVabc * Tabc (:257); the algorithm reports its real part (:76).
PARITY: pinned by construction checks in tests/test_complex_triples.py -- with real inputs it must
reproduce the real oracle (oracle/pt_oracle.py, itself pinned by the reference's recorded UEG (T) energy);
the reference holds no known answer for complex inputs ("parity unpinned" for the imaginary parts).

Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

PERM = [(0, 1, 2), (1, 0, 2), (1, 2, 0), (0, 2, 1), (2, 0, 1), (2, 1, 0)]    # Permutation<3>(p).images
SF = {3: 8.0, 1: -4.0, 0: 2.0}                                              # by invariantElementsCount()


def _invariant(s):
    return sum(1 for m in range(3) if PERM[s][m] == m)


def _permuted(W, letters):
    """W[letters] read as an "abc"-indexed array: result[a,b,c] = W[x_letters0, x_letters1, x_letters2]."""
    return np.einsum(f"{letters}->abc", W)


def _str_perm(s, p):
    return "".join(s[PERM[p][m]] for m in range(3))


def particle_hole_block(T2, Vphhh, conj_gab, gai, x, y, z):
    """DVabc for the hole order (x,y,z) (:341-348 + :135-140)."""
    W = np.einsum("ad,Fdb,Fc->abc", T2[:, :, x, y], conj_gab, gai[:, :, z], optimize=True)
    W -= np.einsum("abl,cl->abc", T2[:, :, x, :], Vphhh[:, :, z, y], optimize=True)
    return W


def triples_complex(epsi, epsa, T1, T2, Vpphh, Vphhh, gamma, return_per_triple=False):
    """E(T) (complex) of calculate() (:166-271).  T1[v,o], T2[v,v,o,o], Vpphh[v,v,o,o], Vphhh[v,o,o,o],
    gamma[NF,Np,Np], all complex (real arrays are accepted); particles are the last v states (:44-45)."""
    o, v = len(epsi), len(epsa)
    np_ = gamma.shape[1]
    a0 = np_ - v
    conj_gab = np.conj(gamma[:, a0:, a0:])        # conjGammaFab (:312-318)
    gai = gamma[:, a0:, :o]
    e_tot = 0.0 + 0.0j
    per = []
    for i in range(o):
        for j in range(i, o):
            for k in range(j, o):
                h = (i, j, k)
                piDV, distinct, hp = [None] * 6, [False] * 6, [None] * 6
                DV = np.zeros((v, v, v), dtype=complex)
                for p in range(6):
                    hp[p] = tuple(h[PERM[p][m]] for m in range(3))
                    q = next((q for q in range(p) if hp[q] == hp[p]), p)
                    if q < p:
                        piDV[p] = piDV[q]
                    else:
                        distinct[p] = True
                        piDV[p] = particle_hole_block(T2, Vphhh, conj_gab, gai, *hp[p])
                    DV += _permuted(piDV[p], _str_perm("abc", p))                      # :221
                D = (epsi[i] + epsi[j] + epsi[k] - epsa[:, None, None] - epsa[None, :, None] - epsa[None, None, :])
                DV = np.conj(DV / D)                                                  # :224-230
                e = 0.0 + 0.0j
                for p in range(6):
                    if not distinct[p]:
                        continue
                    x, y, z = hp[p]
                    SV = 0.5 * np.einsum("a,bc->abc", T1[:, x], Vpphh[:, :, y, z])    # :142-148
                    Tabc = np.zeros((v, v, v), dtype=complex)
                    for s in range(6):
                        letters = _str_perm(_str_perm("abc", s), p)                   # "abc" * sigma * pi
                        sf = SF[_invariant(s)]
                        Tabc += sf * _permuted(piDV[p], letters) + sf * _permuted(SV, letters)
                    e += np.sum(DV * Tabc)                                            # :257
                per.append(e)
                e_tot += e
    return (e_tot, np.array(per)) if return_per_triple else e_tot
