"""Complex closed-shell perturbative triples on the device -- SURVEY.md 8f, N4 (first slice).

Reference: CcsdPerturbativeTriplesComplex::Calculator<complex>::calculate
(src/algorithms/CcsdPerturbativeTriplesComplex.cxx:166-271): complex amplitudes and integrals, the
particle term from conj(GammaFab)["Fdb"] GammaFai["Fck"] (:341-348), the hole term from
PHHHCoulombIntegrals["clkj"] (:135-140), DV <- conj(DV / Delta) (:224-230), real part of the energy (:76).

The step reuses the real step's machinery (fused DMMA kernel, item / orbit tables, TMEM accumulation,
fused epilogue) through an exact rewriting in real arithmetic.  With W = W_r + i W_i the triples block
of one hole permutation, S likewise the singles term, and the fused epilogue's bilinear form
B(W; S) = sum_x (Xd + Sd)[x] (sum_nu c_nu Xd[x o nu]) / D[x],

    Re E(T) = B(W_r; S_r) + B(W_i; S_i),

because conj() only flips the sign of the imaginary part of one factor and the permutation / spin-factor
algebra is real.  Each of the two passes is a REAL (T) evaluation whose contractions have their real and
imaginary parts stacked along the contracted index:

    W_r = [T_r | -T_i] . [V_r ; V_i]   (particle term, contraction length 2v; hole term 2o alike)
    W_i = [T_i |  T_r] . [V_r ; V_i]
    S_r = 1/2 (t_r (x) P_r - t_i (x) P_i),   S_i = 1/2 (t_r (x) P_i + t_i (x) P_r)   (two singles terms)

so the big stacked integrals [V_r ; V_i] are packed once and shared by both passes.  Cost: 4x the real
step, as complex arithmetic demands.  The complex PPPH block V[b,c,d,k] = sum_F conj(G[F,d,b]) G[F,c,k]
is built on the device by the tensor engine.
"""
from __future__ import annotations

import numpy as np

from .tensor_engine import DeviceTensors
from .triples import Algorithm, SisiException, TriplesEngine, register


def complex_ppph_from_vertex(gamma, o, v, device=0):
    """(V_r, V_i)[b,c,d,k] of V = conj(GammaFab)["Fdb"] GammaFai["Fck"] (:341-348), column-major [v,v,v,o]."""
    np_ = gamma.shape[1]
    a0 = np_ - v
    gab, gai = gamma[:, a0:, a0:], gamma[:, a0:, :o]
    with DeviceTensors(device) as eng:
        abr, abi = eng.tensor(gab.shape, gab.real), eng.tensor(gab.shape, gab.imag)
        air, aii = eng.tensor(gai.shape, gai.real), eng.tensor(gai.shape, gai.imag)
        vr, vi = eng.tensor((v, v, v, o)), eng.tensor((v, v, v, o))
        eng.contract(1.0, abr, "Fdb", air, "Fck", 0.0, vr, "bcdk")      # Re: ab_r ai_r + ab_i ai_i
        eng.contract(1.0, abi, "Fdb", aii, "Fck", 1.0, vr, "bcdk")
        eng.contract(1.0, abr, "Fdb", aii, "Fck", 0.0, vi, "bcdk")      # Im: ab_r ai_i - ab_i ai_r
        eng.contract(-1.0, abi, "Fdb", air, "Fck", 1.0, vi, "bcdk")
        return vr.get(), vi.get()


def complex_triples_energy(epsi, epsa, T1, T2, Vpphh, Vphhh, gamma, device: int = 0, return_per_triple=False):
    """Re E(T) of the complex closed-shell step.  T1[v,o], T2[v,v,o,o], Vpphh[v,v,o,o], Vphhh[v,o,o,o],
    gamma[NF,Np,Np] (complex or real arrays, column-major index order of the reference)."""
    o, v = int(len(epsi)), int(len(epsa))
    c = lambda a: np.asarray(a, dtype=np.complex128)
    T1, T2, P, U = c(T1), c(T2), c(Vpphh), np.einsum("clzy->yzlc", c(Vphhh))     # U[y,z,l,c] = Vphhh[c,l,z,y]
    vr, vi = complex_ppph_from_vertex(np.asarray(gamma), o, v, device)
    f = np.asfortranarray
    ppph_e = f(np.concatenate([vr, vi], axis=2))                  # [v,v,2v,o]: [V_r ; V_i] along d
    hhhp_e = f(np.concatenate([U.real, U.imag], axis=2))          # [o,o,2o,v]: [U_r ; U_i] along l
    total, per = 0.0, None
    with TriplesEngine(o, v, device=device, o_all=2 * o, vd=2 * v) as eng:
        eng.set_eigenenergies(epsi, epsa)
        eng.set_hhhp(hhhp_e)
        eng.set_ppph(ppph_e)
        for part in ("re", "im"):
            if part == "re":    # W_r, S_r
                ta, tb = T2.real, -T2.imag
                s1, s2 = (T1.real, P.real), (-T1.imag, P.imag)
            else:               # W_i, S_i
                ta, tb = T2.imag, T2.real
                s1, s2 = (T1.real, P.imag), (T1.imag, P.real)
            eng.set_doubles(f(np.concatenate([ta, tb], axis=1)))          # [v,2v,o,o]
            eng.set_doubles_hole(f(np.concatenate([ta, tb], axis=3)))     # [v,v,o,2o]
            eng.set_singles(f(s1[0])); eng.set_pphh(f(s1[1]))
            eng.set_singles_pair(f(s2[0]), f(s2[1]))
            res = eng.run()
            total += res.energy
            per = res.per_triple if per is None else per + res.per_triple
        stats = eng.stats()
    return (total, per, stats) if return_per_triple else total


@register
class CcsdPerturbativeTriplesComplex(Algorithm):
    """Plan step under the reference's name and argument keys (CcsdPerturbativeTriplesComplex.cxx:32-84):
    in: HoleEigenEnergies, ParticleEigenEnergies, CoulombVertex, CcsdSinglesAmplitudes,
    CcsdDoublesAmplitudes, PPHHCoulombIntegrals, PHHHCoulombIntegrals, CcsdEnergy;
    out: CcsdPerturbativeTriplesComplexEnergy = CcsdEnergy + Re E(T)."""
    name = "CcsdPerturbativeTriplesComplex"

    def run(self):
        e_t = complex_triples_energy(
            self.getTensorArgument("HoleEigenEnergies"), self.getTensorArgument("ParticleEigenEnergies"),
            self.getTensorArgument("CcsdSinglesAmplitudes"), self.getTensorArgument("CcsdDoublesAmplitudes"),
            self.getTensorArgument("PPHHCoulombIntegrals"), self.getTensorArgument("PHHHCoulombIntegrals"),
            self.getTensorArgument("CoulombVertex"), device=self.getIntegerArgument("device", 0))
        e_ccsd = self.getRealArgument("CcsdEnergy")                                     # :78
        e = e_ccsd + e_t
        self.log = {"e": e, "ccsd": e_ccsd, "triples": e_t}
        if not self.isArgumentGiven("CcsdPerturbativeTriplesComplexEnergy"):
            raise SisiException("Missing argument: CcsdPerturbativeTriplesComplexEnergy")
        self.setRealArgument("CcsdPerturbativeTriplesComplexEnergy", e)                 # :82
        return e
