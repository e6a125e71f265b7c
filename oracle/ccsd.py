"""Closed-shell CCSD amplitudes for SMALL systems -- TEST INFRASTRUCTURE (oracle side).

The (T) step consumes converged CCSD amplitudes (reference producers:
CcsdEnergyFromCoulombIntegralsReference.cxx:29-295 driven by ClusterSinglesDoublesAlgorithm.cxx:37-128
with the DiisMixer).  To pin the (T) restatement against the known answers the reference holds
(integration-tests/tests/cc4s/*/cc4s.correct.out.yaml) converged amplitudes are needed, so this module
solves the CCSD equations in plain NumPy.  It does NOT follow the reference's closed-shell residuum
line by line: it uses the textbook spin-orbital form (Stanton, Gauss, Watts, Bartlett, J. Chem. Phys.
94, 4334 (1991)) with the canonical Hartree-Fock assumption of the reference (Fock matrix =
diag(eigenenergies), ClusterSinglesDoublesAlgorithm.cxx:302-331); the converged solution is the same,
which the CCSD correlation energy test checks against the reference's value.

Inputs: eigenenergies and the Coulomb vertex Gamma[F,p,q] (holes first), integrals
V_pqrs = <pq|rs> = Re.Re + Im.Im of Gamma[F,p,r], Gamma[F,q,s] (CoulombIntegralsFromVertex.cxx:399-433).
"""
from __future__ import annotations

import numpy as np


def spatial_integrals(gamma: np.ndarray) -> np.ndarray:
    """V[p,q,r,s] = sum_F Re G[F,p,r] Re G[F,q,s] + Im G[F,p,r] Im G[F,q,s]."""
    gr, gi = gamma.real, gamma.imag
    return np.einsum("fpr,fqs->pqrs", gr, gr, optimize=True) + np.einsum("fpr,fqs->pqrs", gi, gi, optimize=True)


def solve(epsi, epsa, gamma, tol=1e-11, max_iter=100, diis=6, log=None):
    """Returns dict(T1[v,o], T2[v,v,o,o] spatial closed-shell amplitudes, energy, iterations)."""
    no, nv = len(epsi), len(epsa)
    n = no + nv
    V = spatial_integrals(gamma)                                  # <pq|rs>, spatial
    # spin orbitals: index = spin * n + spatial; occupied first in the o/v lists below
    spin = np.repeat([0, 1], n)
    spat = np.tile(np.arange(n), 2)
    occ = np.concatenate([np.arange(no), n + np.arange(no)])
    vir = np.concatenate([no + np.arange(nv), n + no + np.arange(nv)])
    order = np.concatenate([occ, vir])
    sp, sa = spin[order], spat[order]
    N, O = 2 * n, 2 * no
    same = (sp[:, None] == sp[None, :]).astype(np.float64)
    Vs = V[np.ix_(sa, sa, sa, sa)] * same[:, None, :, None] * same[None, :, None, :]     # <PQ|RS>
    A = Vs - Vs.transpose(0, 1, 3, 2)                                                  # <PQ||RS>
    del Vs
    eps = np.concatenate([epsi, epsa])[sa]
    o, v = slice(0, O), slice(O, N)
    eo, ev = eps[o], eps[v]
    D1 = eo[:, None] - ev[None, :]
    D2 = eo[:, None, None, None] + eo[None, :, None, None] - ev[None, None, :, None] - ev[None, None, None, :]
    foo, fvv = np.diag(eo), np.diag(ev)
    t1 = np.zeros((O, N - O))
    t2 = A[o, o, v, v] / D2
    es = lambda *a: np.einsum(*a, optimize=True)

    def energy(t1, t2):
        return 0.25 * es("ijab,ijab->", A[o, o, v, v], t2) + 0.5 * es("ijab,ia,jb->", A[o, o, v, v], t1, t1)

    hist_t, hist_e = [], []
    e_old = energy(t1, t2)
    for it in range(1, max_iter + 1):
        tt = es("ia,jb->ijab", t1, t1)
        tau_t = t2 + 0.5 * (tt - tt.transpose(0, 1, 3, 2))
        tau = t2 + tt - tt.transpose(0, 1, 3, 2)
        Fae = fvv - np.diag(np.diag(fvv)) + es("mf,mafe->ae", t1, A[o, v, v, v]) - 0.5 * es("mnaf,mnef->ae", tau_t, A[o, o, v, v])
        Fmi = foo - np.diag(np.diag(foo)) + es("ne,mnie->mi", t1, A[o, o, o, v]) + 0.5 * es("inef,mnef->mi", tau_t, A[o, o, v, v])
        Fme = es("nf,mnef->me", t1, A[o, o, v, v])
        Wmnij = A[o, o, o, o] + es("je,mnie->mnij", t1, A[o, o, o, v]) - es("ie,mnje->mnij", t1, A[o, o, o, v]) \
            + 0.25 * es("ijef,mnef->mnij", tau, A[o, o, v, v])
        Wabef = A[v, v, v, v] - es("mb,amef->abef", t1, A[v, o, v, v]) + es("ma,bmef->abef", t1, A[v, o, v, v]) \
            + 0.25 * es("mnab,mnef->abef", tau, A[o, o, v, v])
        Wmbej = A[o, v, v, o] + es("jf,mbef->mbej", t1, A[o, v, v, v]) - es("nb,mnej->mbej", t1, A[o, o, v, o]) \
            - es("jnfb,mnef->mbej", 0.5 * t2 + es("jf,nb->jnfb", t1, t1), A[o, o, v, v])
        r1 = es("ie,ae->ia", t1, Fae) - es("ma,mi->ia", t1, Fmi) + es("imae,me->ia", t2, Fme) \
            - es("nf,naif->ia", t1, A[o, v, o, v]) - 0.5 * es("imef,maef->ia", t2, A[o, v, v, v]) \
            - 0.5 * es("mnae,nmei->ia", t2, A[o, o, v, o])
        r2 = A[o, o, v, v].copy()
        x = es("ijae,be->ijab", t2, Fae - 0.5 * es("mb,me->be", t1, Fme))
        r2 += x - x.transpose(0, 1, 3, 2)
        x = es("imab,mj->ijab", t2, Fmi + 0.5 * es("je,me->mj", t1, Fme))
        r2 -= x - x.transpose(1, 0, 2, 3)
        r2 += 0.5 * es("mnab,mnij->ijab", tau, Wmnij) + 0.5 * es("ijef,abef->ijab", tau, Wabef)
        x = es("imae,mbej->ijab", t2, Wmbej) - es("ie,ma,mbej->ijab", t1, t1, A[o, v, v, o])
        r2 += x - x.transpose(1, 0, 2, 3) - x.transpose(0, 1, 3, 2) + x.transpose(1, 0, 3, 2)
        x = es("ie,abej->ijab", t1, A[v, v, v, o])
        r2 += x - x.transpose(1, 0, 2, 3)
        x = es("ma,mbij->ijab", t1, A[o, v, o, o])
        r2 -= x - x.transpose(0, 1, 3, 2)
        n1, n2 = r1 / D1, r2 / D2
        err = np.concatenate([(n1 - t1).ravel(), (n2 - t2).ravel()])
        t1, t2 = n1, n2
        hist_t.append(np.concatenate([t1.ravel(), t2.ravel()]))
        hist_e.append(err)
        if len(hist_t) > diis:
            hist_t.pop(0); hist_e.pop(0)
        if len(hist_t) > 1:
            m = len(hist_t)
            B = -np.ones((m + 1, m + 1)); B[m, m] = 0.0
            for a in range(m):
                for b in range(m):
                    B[a, b] = hist_e[a] @ hist_e[b]
            rhs = np.zeros(m + 1); rhs[m] = -1.0
            c = np.linalg.lstsq(B, rhs, rcond=None)[0][:m]
            mix = sum(ci * ti for ci, ti in zip(c, hist_t))
            t1 = mix[:t1.size].reshape(t1.shape)
            t2 = mix[t1.size:].reshape(t2.shape)
        e = energy(t1, t2)
        rn = float(np.abs(err).max())
        if log:
            log(f"ccsd iter {it:3d}  E = {e:.15f}  dE = {e - e_old:+.2e}  |dt|max = {rn:.2e}")
        if abs(e - e_old) < tol and rn < 10 * tol ** 0.5 * 1e-2:
            break
        e_old = e
    # closed-shell spatial amplitudes: T2[a,b,i,j] = t(i alpha, j beta -> a alpha, b beta)
    ia, ib = np.arange(no), no + np.arange(no)           # occupied alpha / beta positions in `order`
    aa, ab = np.arange(nv), nv + np.arange(nv)           # virtual alpha / beta positions
    T2 = t2[np.ix_(ia, ib, aa, ab)].transpose(2, 3, 0, 1)
    T1 = t1[np.ix_(ia, aa)].T
    return {"T1": np.asfortranarray(T1), "T2": np.asfortranarray(T2), "energy": float(e), "iterations": it}
