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
is built on the device by the tensor engine.  The driver itself lives in the library
(`pt_complex_triples`, csrc/pt_complex.cu); this module binds it and registers the plan step.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .triples import Algorithm, SisiException, register


def complex_triples_energy(epsi, epsa, T1, T2, Vpphh, Vphhh, gamma, device: int = 0, return_per_triple=False):
    """Re E(T) of the complex closed-shell step through the C ABI (pt_complex_triples, csrc/pt_complex.cu).
    T1[v,o], T2[v,v,o,o], Vpphh[v,v,o,o], Vphhh[v,o,o,o], gamma[NF,Np,Np] (complex or real arrays,
    column-major index order of the reference)."""
    lib = _lib.load()
    o, v = int(len(epsi)), int(len(epsa))
    f = lambda a: np.asfortranarray(a, dtype=np.float64)
    parts = []
    for a, shape, name in ((T1, (v, o), "CcsdSinglesAmplitudes"), (T2, (v, v, o, o), "CcsdDoublesAmplitudes"),
                           (Vpphh, (v, v, o, o), "PPHHCoulombIntegrals"), (Vphhh, (v, o, o, o), "PHHHCoulombIntegrals")):
        a = np.asarray(a)
        if tuple(a.shape) != shape:
            raise ValueError(f"{name}: expected shape {shape}, got {tuple(a.shape)}")
        parts += [f(a.real), f(a.imag) if np.iscomplexobj(a) else np.zeros(shape, order="F")]
    g = np.asarray(gamma)
    nf, np_, np2 = g.shape
    if np_ != np2 or np_ < o + v:
        raise ValueError("CoulombVertex must be [NF,Np,Np] with Np >= No + Nv")
    gre, gim = f(g.real), (f(g.imag) if np.iscomplexobj(g) else np.zeros(g.shape, order="F"))
    ei, ea = f(epsi), f(epsa)
    per = np.zeros(o * (o + 1) * (o + 2) // 6)
    e = C.c_double(0.0)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    _lib.check(lib.pt_complex_triples(o, v, int(device), p(ei), p(ea), *[p(a) for a in parts], nf, np_, p(gre), p(gim),
                                      C.byref(e), p(per)))
    return (float(e.value), per) if return_per_triple else float(e.value)


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
