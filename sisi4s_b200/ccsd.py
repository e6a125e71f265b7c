"""CCSD amplitude solver for the step BEFORE the (T) path (first slice).

Produces what the reference's CcsdEnergyFromCoulombIntegralsReference (src/algorithms/
CcsdEnergyFromCoulombIntegralsReference.cxx:29-295, iteration loop ClusterSinglesDoublesAlgorithm.cxx:
37-128, DIIS src/mixers/DiisMixer.cxx:103-181) hands to the triples step: CcsdEnergy,
CcsdSinglesAmplitudes[v,o], CcsdDoublesAmplitudes[v,v,o,o], converged to the same thresholds
(energyConvergence / amplitudesConvergence).  The contractions are library GEMMs (torch.einsum in
FP64 on the device of the inputs -> cuBLAS / cuTENSOR on a GPU); there is no hand-written kernel here,
and the equations are the textbook spin-orbital form with a canonical Hartree-Fock reference
(Fock = diag(eigenenergies), as the reference assumes), not a transcription of the reference's
closed-shell residuum.  The converged solution is the same: tests/test_ccsd_step.py checks the CCSD
energy against the value the reference records for its UEG test system.

Memory is O((2v)^4): meant for the small and medium systems of tests and examples (v <~ 120 on one GPU).
"""
from __future__ import annotations

import torch


def _spin_blocks(no, nv, device):
    n = no + nv
    spin = torch.arange(2, device=device).repeat_interleave(n)
    spat = torch.arange(n, device=device).repeat(2)
    occ = torch.cat([torch.arange(no, device=device), n + torch.arange(no, device=device)])
    vir = torch.cat([no + torch.arange(nv, device=device), n + no + torch.arange(nv, device=device)])
    order = torch.cat([occ, vir])
    return spin[order], spat[order]


def solve_ccsd(epsi, epsa, gamma, device, energy_convergence=1e-8, amplitudes_convergence=1e-8,
               max_iterations=50, max_residua=4, log=None):
    """gamma: complex vertex [NF, Np, Np] (holes first).  Returns dict(energy, T1[v,o], T2[v,v,o,o],
    iterations, converged) with CPU numpy amplitudes in the reference's index order."""
    dev = torch.device(device)
    f64 = torch.float64
    epsi = torch.as_tensor(epsi, dtype=f64, device=dev)
    epsa = torch.as_tensor(epsa, dtype=f64, device=dev)
    g = torch.as_tensor(gamma, device=dev)
    gr, gi = g.real.to(f64).contiguous(), g.imag.to(f64).contiguous()
    no, nv = epsi.numel(), epsa.numel()
    es = torch.einsum
    # <pq|rs> = Re.Re + Im.Im of G[F,p,r], G[F,q,s]  (CoulombIntegralsFromVertex.cxx:399-433)
    V = es("fpr,fqs->pqrs", gr, gr) + es("fpr,fqs->pqrs", gi, gi)
    sp, sa = _spin_blocks(no, nv, dev)
    same = (sp[:, None] == sp[None, :]).to(f64)
    Vs = V[sa][:, sa][:, :, sa][:, :, :, sa] * same[:, None, :, None] * same[None, :, None, :]
    A = Vs - Vs.permute(0, 1, 3, 2)
    del V, Vs
    N, O = 2 * (no + nv), 2 * no
    eps = torch.cat([epsi, epsa])[sa]
    o, v = slice(0, O), slice(O, N)
    eo, ev = eps[o], eps[v]
    D1 = eo[:, None] - ev[None, :]
    D2 = eo[:, None, None, None] + eo[None, :, None, None] - ev[None, None, :, None] - ev[None, None, None, :]
    Aoovv, Aooov, Aovvv = A[o, o, v, v].contiguous(), A[o, o, o, v].contiguous(), A[o, v, v, v].contiguous()
    Aoooo, Avvvv, Aovvo = A[o, o, o, o].contiguous(), A[v, v, v, v].contiguous(), A[o, v, v, o].contiguous()
    Avovv, Aoovo, Aovov = A[v, o, v, v].contiguous(), A[o, o, v, o].contiguous(), A[o, v, o, v].contiguous()
    Avvvo, Aovoo = A[v, v, v, o].contiguous(), A[o, v, o, o].contiguous()
    del A
    t1 = torch.zeros((O, N - O), dtype=f64, device=dev)
    t2 = Aoovv / D2

    def energy(t1, t2):
        return float(0.25 * es("ijab,ijab->", Aoovv, t2) + 0.5 * es("ijab,ia,jb->", Aoovv, t1, t1))

    hist_t, hist_e = [], []
    e_old = energy(t1, t2)
    converged = False
    for it in range(1, max_iterations + 1):
        tt = es("ia,jb->ijab", t1, t1)
        tau_t = t2 + 0.5 * (tt - tt.permute(0, 1, 3, 2))
        tau = t2 + tt - tt.permute(0, 1, 3, 2)
        Fae = es("mf,mafe->ae", t1, Aovvv) - 0.5 * es("mnaf,mnef->ae", tau_t, Aoovv)
        Fmi = es("ne,mnie->mi", t1, Aooov) + 0.5 * es("inef,mnef->mi", tau_t, Aoovv)
        Fme = es("nf,mnef->me", t1, Aoovv)
        Wmnij = Aoooo + es("je,mnie->mnij", t1, Aooov) - es("ie,mnje->mnij", t1, Aooov) \
            + 0.25 * es("ijef,mnef->mnij", tau, Aoovv)
        Wabef = Avvvv - es("mb,amef->abef", t1, Avovv) + es("ma,bmef->abef", t1, Avovv) \
            + 0.25 * es("mnab,mnef->abef", tau, Aoovv)
        Wmbej = Aovvo + es("jf,mbef->mbej", t1, Aovvv) - es("nb,mnej->mbej", t1, Aoovo) \
            - es("jnfb,mnef->mbej", 0.5 * t2 + es("jf,nb->jnfb", t1, t1), Aoovv)
        r1 = es("ie,ae->ia", t1, Fae) - es("ma,mi->ia", t1, Fmi) + es("imae,me->ia", t2, Fme) \
            - es("nf,naif->ia", t1, Aovov) - 0.5 * es("imef,maef->ia", t2, Aovvv) \
            - 0.5 * es("mnae,nmei->ia", t2, Aoovo)
        r2 = Aoovv.clone()
        x = es("ijae,be->ijab", t2, Fae - 0.5 * es("mb,me->be", t1, Fme))
        r2 += x - x.permute(0, 1, 3, 2)
        x = es("imab,mj->ijab", t2, Fmi + 0.5 * es("je,me->mj", t1, Fme))
        r2 -= x - x.permute(1, 0, 2, 3)
        r2 += 0.5 * es("mnab,mnij->ijab", tau, Wmnij) + 0.5 * es("ijef,abef->ijab", tau, Wabef)
        x = es("imae,mbej->ijab", t2, Wmbej) - es("ie,ma,mbej->ijab", t1, t1, Aovvo)
        r2 += x - x.permute(1, 0, 2, 3) - x.permute(0, 1, 3, 2) + x.permute(1, 0, 3, 2)
        x = es("ie,abej->ijab", t1, Avvvo)
        r2 += x - x.permute(1, 0, 2, 3)
        x = es("ma,mbij->ijab", t1, Aovoo)
        r2 -= x - x.permute(0, 1, 3, 2)
        n1, n2 = r1 / D1, r2 / D2
        err = torch.cat([(n1 - t1).reshape(-1), (n2 - t2).reshape(-1)])
        t1, t2 = n1, n2
        # DIIS over the last max_residua iterates (DiisMixer.cxx:103-181)
        hist_t.append(torch.cat([t1.reshape(-1), t2.reshape(-1)]))
        hist_e.append(err)
        if len(hist_t) > max_residua:
            hist_t.pop(0); hist_e.pop(0)
        m = len(hist_t)
        if m > 1:
            B = -torch.ones((m + 1, m + 1), dtype=f64, device=dev)
            B[m, m] = 0.0
            E = torch.stack(hist_e)
            B[:m, :m] = E @ E.T
            rhs = torch.zeros(m + 1, dtype=f64, device=dev)
            rhs[m] = -1.0
            c = torch.linalg.lstsq(B, rhs[:, None]).solution[:m, 0]
            mix = (c[:, None] * torch.stack(hist_t)).sum(0)
            t1 = mix[:t1.numel()].reshape(t1.shape)
            t2 = mix[t1.numel():].reshape(t2.shape)
        e = energy(t1, t2)
        dt = float(err.abs().max())
        if log:
            log(f"  iteration={it} energy={e:.15g} dE={e - e_old:+.2e} |dT|max={dt:.2e}")
        if abs(e - e_old) < energy_convergence and dt < amplitudes_convergence:
            converged = True
            break
        e_old = e
    ia, ib = torch.arange(no, device=dev), no + torch.arange(no, device=dev)
    aa, ab = torch.arange(nv, device=dev), nv + torch.arange(nv, device=dev)
    T2 = t2[ia][:, ib][:, :, aa][:, :, :, ab].permute(2, 3, 0, 1)
    T1 = t1[ia][:, aa].T
    import numpy as np
    return {"energy": e, "iterations": it, "converged": converged,
            "T1": np.asfortranarray(T1.cpu().numpy()), "T2": np.asfortranarray(T2.cpu().numpy())}
