"""Uniform-electron-gas inputs for the (T) step: host-side input generator (like synthetic.py),
the counterpart of the reference's UegVertexGenerator algorithm (registered under that name in
plan.py).  No part of the (T) computation happens here.

Restates the model the reference generates in UegVertexGenerator::run (reference
src/algorithms/UegVertexGenerator.cxx:51-229; Madelung constant :10-30, exchange :37-41, plane-wave
grid :66-91, cell volume :94-96, Hartree-Fock eigenenergies :104-114, momentum-transfer grid
:122-146, vertex elements sqrt(4 pi / (V G^2)) with the Madelung term at G = 0 :217-226).

The reference's generator writes the vertex in the plane-wave basis "just for profiling" (:148-149):
with complex orbitals the real-integral formulas of CoulombIntegralsFromVertex.cxx:399-433
(V = Re.Re + Im.Im) do not conserve momentum.  Here the degenerate +k/-k plane waves are rotated to
REAL standing waves (cos, sin), for which Gamma^G_pq = Gamma^G_qp and Re.Re + Im.Im is exactly the
Coulomb integral.  Eigenenergies and all correlation energies are invariant under that rotation, so
the known answers the reference holds for this system
(integration-tests/tests/cc4s/ueg/rs1.0-7occ-26virt/cc4s.correct.out.yaml:124-169: MP2, CCSD, (T))
apply unchanged.
"""
from __future__ import annotations

import math

import numpy as np


def madelung(v: float) -> float:
    """UegVertexGenerator::evalMadelung (:10-30)."""
    kappa = v ** (-1.0 / 3.0)
    term2 = math.pi / (kappa * kappa * v)
    term4 = 2 * kappa / math.sqrt(math.pi)
    box = 1.0 / kappa
    recipsum = realsum = 0.0
    for l1 in range(-6, 7):
        for l2 in range(-6, 7):
            for l3 in range(-6, 7):
                n2 = l1 * l1 + l2 * l2 + l3 * l3
                if n2 > 0:
                    modr = box * math.sqrt(n2)
                    k2 = kappa * kappa * n2
                    recipsum -= 1.0 / (math.pi * k2) * math.exp(-math.pi * math.pi * k2 / kappa / kappa) / v
                    realsum -= math.erfc(kappa * modr) / modr
    return realsum + term2 + term4 + recipsum


def plane_wave_grid(no: int, nv: int):
    """Integer k-grid sorted by length, first no+nv points (:66-91); closed shells are required."""
    n = no + nv
    mg = int((5.0 * n) ** (1.0 / 3.0))
    pts = [(a, b, c) for a in range(-mg, mg + 1) for b in range(-mg, mg + 1) for c in range(-mg, mg + 1)]
    pts.sort(key=lambda t: t[0] * t[0] + t[1] * t[1] + t[2] * t[2])     # stable, like std::sort on equal keys is not
    sl = lambda t: t[0] * t[0] + t[1] * t[1] + t[2] * t[2]
    if sl(pts[no]) == sl(pts[no - 1]) or sl(pts[n]) == sl(pts[n - 1]):
        raise ValueError("occupied / virtual orbitals do not form closed shells")
    return np.array(pts[:n], dtype=np.int64)


def make_ueg(no: int, nv: int, rs: float):
    """Returns (epsi[no], epsa[nv], Gamma[NF, Np, Np] complex in the real standing-wave basis)."""
    n = no + nv
    k = plane_wave_grid(no, nv)
    vol = rs ** 3 / 3.0 * 4.0 * math.pi * no * 2                         # :94
    b = 2.0 * math.pi / vol ** (1.0 / 3.0)
    mad = madelung(vol)
    kd = b * k.astype(np.float64)
    # Hartree-Fock eigenenergies (:104-109): kinetic - exchange with the occupied states
    eps = np.empty(n)
    for p in range(n):
        ex = 0.0
        for o in range(no):
            q2 = float(((kd[p] - kd[o]) ** 2).sum())
            ex += mad if q2 < 1e-8 else 4.0 * math.pi / vol / q2
        eps[p] = 0.5 * float((kd[p] ** 2).sum()) - ex
    # momentum-transfer grid (:122-146)
    diff = k[:, None, :] - k[None, :, :]
    max_r = int((diff ** 2).sum(-1).max())
    mg = int(np.abs(diff).max())
    mom = {}
    for g1 in range(-mg, mg + 1):
        for g2 in range(-mg, mg + 1):
            for g3 in range(-mg, mg + 1):
                if g1 * g1 + g2 * g2 + g3 * g3 <= max_r:
                    mom[(g1, g2, g3)] = len(mom)
    nf = len(mom)
    # plane-wave vertex G[F, q, p] = sqrt(w) for F = k_q - k_p (:217-226)
    gpw = np.zeros((nf, n, n))
    fac = 4.0 * math.pi / vol
    for q in range(n):
        for p in range(n):
            d = tuple(int(x) for x in (k[q] - k[p]))
            s = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
            gpw[mom[d], q, p] = math.sqrt(fac / (b * b * s) if s else mad)
    # rotation of each degenerate (+k, -k) pair to real standing waves
    index = {tuple(int(x) for x in kk): m for m, kk in enumerate(k)}
    C = np.zeros((n, n), dtype=np.complex128)                            # |chi_m> = sum_k C[k, m] |k>
    done = set()
    for m, kk in enumerate(k):
        if m in done:
            continue
        mm = index[tuple(int(-x) for x in kk)]
        if mm == m:
            C[m, m] = 1.0
        else:
            if (m < no) != (mm < no):
                raise ValueError("+k and -k on different sides of the Fermi level")
            C[m, m] = C[mm, m] = 1.0 / math.sqrt(2.0)                    # cos
            C[m, mm] = -1j / math.sqrt(2.0)                              # sin
            C[mm, mm] = 1j / math.sqrt(2.0)
            done.add(mm)
        done.add(m)
    # Gamma'[F, mu, nu] = sum_{q p} conj(C[q, mu]) G[F, q, p] C[p, nu]
    gamma = np.einsum("qm,fqp,pn->fmn", C.conj(), gpw.astype(np.complex128), C, optimize=True)
    return eps[:no].copy(), eps[no:].copy(), np.asfortranarray(gamma)


def mp2_energy(epsi, epsa, vpphh) -> float:
    """Closed-shell MP2, E = sum (2 V_abij - V_abji) V_abij / (e_i + e_j - e_a - e_b)."""
    d = epsi[None, None, :, None] + epsi[None, None, None, :] - epsa[:, None, None, None] - epsa[None, :, None, None]
    return float(np.einsum("abij,abij->", 2.0 * vpphh - vpphh.transpose(0, 1, 3, 2), vpphh / d))
