"""Closed-shell CCSD amplitude solver on the device -- the step in front of the (T) path.

Follows the reference statement by statement:

  * residuum: CcsdEnergyFromCoulombIntegralsReference::getResiduum (reference
    src/algorithms/CcsdEnergyFromCoulombIntegralsReference.cxx:29-295) -- every CTF index-string
    statement there is one `contract` / `add` below with the same strings; products of three tensors
    (V * Tai * Tai) are evaluated pairwise through the intermediates named Y../Z.. or through
    Xabij = Tabij + Tai Tbj (the tensor the reference itself builds at :73-74), which merges the
    `... * Tabij` and `... * Tai * Tai` statements that differ only in that factor;
  * loop, convergence test and amplitude update: ClusterSinglesDoublesAlgorithm::run (:37-128),
    estimateAmplitudesFromResiduum (:302-331); energy: getEnergy (:130-205, closed shell);
  * mixers: LinearMixer (src/mixers/LinearMixer.cxx:31-49), DiisMixer (src/mixers/DiisMixer.cxx:103-181;
    the (count+1)^2 solve of :16-41 is done on the host, everything else on the device).

All tensors live on the GPU (sisi4s_b200.tensor_engine.DeviceTensors); contractions run on the
library's FP64 tensor-core GEMM.  There is no CPU path.
"""
from __future__ import annotations

import numpy as np

from .tensor_engine import DeviceTensors

DEFAULT_MAX_ITERATIONS = 16          # ClusterSinglesDoublesAlgorithm.hpp
DEFAULT_ENERGY_CONVERGENCE = 1e-6
DEFAULT_AMPLITUDES_CONVERGENCE = 1e-5
DEFAULT_LEVEL_SHIFT = 0.0
BLOCKS = ("PPHH", "PHPH", "HHHH", "HHHP", "PPPH", "PPPP")


class _LinearMixer:
    """LinearMixer.cxx:31-49 on device tensors."""
    def __init__(self, eng, ratio):
        self.eng, self.ratio, self.last = eng, float(ratio), None

    def append(self, A, R):
        if self.last is not None:
            for new, old in zip(A, self.last):   # next = ratio * next + (1 - ratio) * last
                self.eng.add(1.0 - self.ratio, old, _IDX[len(old.shape)], self.ratio, new, _IDX[len(new.shape)])
            for t in self.last:
                t.free()
        for t in R:
            t.free()
        self.last = A

    def get(self):
        return self.last


class _DiisMixer:
    """DiisMixer.cxx:55-181: ring of the last N (amplitudes, residua), B[i,j] = 2 <R_i|R_j> bordered by
    -1, weights = first column of the inverse of its leading (count+1)^2 block."""
    def __init__(self, eng, max_residua):
        N = self.N = int(max_residua)
        self.eng = eng
        self.amplitudes, self.residua = [None] * N, [None] * N
        self.next_index = self.count = 0
        self.B = np.zeros((N + 1, N + 1))
        self.B[0, 1:] = -1.0
        self.B[1:, 0] = -1.0
        self.next = None
        self.weights = None

    def append(self, A, R):
        eng, N, n = self.eng, self.N, self.next_index
        for old in (self.amplitudes[n], self.residua[n]):
            if old is not None:
                for t in old:
                    t.free()
        self.amplitudes[n], self.residua[n] = A, R
        for i in range(N):
            if self.residua[i] is not None:
                ov = 2.0 * sum(eng.dot(x, y) for x, y in zip(self.residua[i], R))
                self.B[n + 1, i + 1] = self.B[i + 1, n + 1] = ov
        if self.count < N:
            self.count += 1
        dim = self.count + 1
        rhs = np.zeros(dim)
        rhs[0] = -1.0
        col = np.linalg.solve(self.B[:dim, :dim], rhs)      # dsysv_ of :16-41
        self.weights = col
        if self.next is not None:
            for t in self.next:
                t.free()
        self.next = [eng.tensor(a.shape) for a in A]
        for j in range(self.count):
            i = (n + N - j) % N
            for t, a in zip(self.next, self.amplitudes[i]):
                eng.add(col[i + 1], a, _IDX[len(a.shape)], 1.0, t, _IDX[len(a.shape)])
        self.next_index = (n + 1) % N

    def get(self):
        return self.next          # owned by the mixer; replaced by the next append


_IDX = {2: "ai", 4: "abij"}


class CcsdSolver:
    """Device-resident integrals + amplitudes of one closed-shell CCSD calculation."""

    def __init__(self, epsi, epsa, integrals: dict, device: int = 0):
        """integrals: the blocks getResiduum reads (:49,136-140) as column-major arrays:
        PPHH[v,v,o,o], PHPH[v,o,v,o], HHHH[o,o,o,o], HHHP[o,o,o,v], PPPH[v,v,v,o], PPPP[v,v,v,v]."""
        self.no, self.nv = int(len(epsi)), int(len(epsa))
        o, v = self.no, self.nv
        missing = [b for b in BLOCKS if b not in integrals]
        if missing:
            raise ValueError("Missing argument: " + ", ".join(m + "CoulombIntegrals" for m in missing))
        self.eng = eng = DeviceTensors(device)
        self.epsi, self.epsa = eng.tensor((o,), epsi), eng.tensor((v,), epsa)
        self.V = {b: eng.tensor(np.shape(integrals[b]), integrals[b]) for b in BLOCKS}
        # Vabij["baij"], the exchange operand of getEnergy (:175-178)
        self.Vx = eng.tensor((v, v, o, o))
        eng.add(1.0, self.V["PPHH"], "baij", 0.0, self.Vx, "abij")
        t = lambda *s: eng.tensor(s)
        # intermediates of the residuum, allocated once
        self.X = t(v, v, o, o)
        self.Kac, self.Mac, self.Lac = t(v, v), t(v, v), t(v, v)
        self.Kki, self.Mki, self.Lki = t(o, o), t(o, o), t(o, o)
        self.Kck, self.Zki = t(v, o), t(o, o)
        self.Y = t(v, o, o, o)
        self.Xakic, self.Xakci = t(v, o, o, v), t(v, o, v, o)
        self.Xklij, self.Xabcd = t(o, o, o, o), t(v, v, v, v)
        self.S = t(v, v, o, o)

    def close(self):
        self.eng.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ reference statements
    def build_x(self, Tai, Tabij):
        """Xabij["abij"] = Tabij["abij"] + Tai["ai"] * Tai["bj"]   (:73-74)"""
        e = self.eng
        e.add(1.0, Tabij, "abij", 0.0, self.X, "abij")
        e.contract(1.0, Tai, "ai", Tai, "bj", 1.0, self.X, "abij")

    def residuum(self, i, Tai, Tabij, Rai, Rabij, initial_doubles_given=False):
        """getResiduum(i, amplitudes) (:29-295) into Rai, Rabij."""
        e, V = self.eng, self.V
        C, A = e.contract, e.add
        Vabij, Vaibj, Vijkl, Vijka, Vabci, Vabcd = (V[b] for b in BLOCKS)
        if i == 0 and not initial_doubles_given:
            A(0.0, Tai, "ai", 0.0, Rai, "ai")
            A(1.0, Vabij, "abij", 0.0, Rabij, "abij")                          # :52-57: MP2 amplitudes
            return
        X, Y = self.X, self.Y
        self.build_x(Tai, Tabij)
        # Kac (:169-173), with Tabij + Tai Tai = X
        C(-2.0, Vabij, "cdkl", X, "adkl", 0.0, self.Kac, "ac")
        C(1.0, Vabij, "dckl", X, "adkl", 1.0, self.Kac, "ac")
        # Lac - Kac (:177-178)
        C(2.0, Vabci, "cdak", Tai, "dk", 0.0, self.Mac, "ac")
        C(-1.0, Vabci, "dcak", Tai, "dk", 1.0, self.Mac, "ac")
        A(1.0, self.Kac, "ac", 0.0, self.Lac, "ac")                            # :176
        A(1.0, self.Mac, "ac", 1.0, self.Lac, "ac")
        # Kki (:181-184)
        C(2.0, Vabij, "cdkl", X, "cdil", 0.0, self.Kki, "ki")
        C(-1.0, Vabij, "dckl", X, "cdil", 1.0, self.Kki, "ki")
        # Lki - Kki (:188-189)
        C(2.0, Vijka, "klic", Tai, "cl", 0.0, self.Mki, "ki")
        C(-1.0, Vijka, "lkic", Tai, "cl", 1.0, self.Mki, "ki")
        A(1.0, self.Kki, "ki", 0.0, self.Lki, "ki")                            # :187
        A(1.0, self.Mki, "ki", 1.0, self.Lki, "ki")
        # :192-201
        C(1.0, self.Lac, "ac", Tabij, "cbij", 0.0, Rabij, "abij")
        C(-1.0, self.Lki, "ki", Tabij, "abkj", 1.0, Rabij, "abij")
        C(1.0, Vabci, "baci", Tai, "cj", 1.0, Rabij, "abij")
        C(1.0, Vaibj, "bkci", Tai, "cj", 0.0, Y, "bkij")                       # :198 = - (Vaibj Tai) Tai
        C(-1.0, Y, "bkij", Tai, "ak", 1.0, Rabij, "abij")
        C(-1.0, Vijka, "jika", Tai, "bk", 1.0, Rabij, "abij")
        C(1.0, Vabij, "acik", Tai, "cj", 0.0, Y, "aikj")                       # :201
        C(-1.0, Y, "aikj", Tai, "bk", 1.0, Rabij, "abij")
        # Xakic (:204-210)
        A(1.0, Vabij, "acik", 0.0, self.Xakic, "akic")
        C(-1.0, Vijka, "lkic", Tai, "al", 1.0, self.Xakic, "akic")
        C(1.0, Vabci, "acdk", Tai, "di", 1.0, self.Xakic, "akic")
        C(-0.5, Vabij, "dclk", Tabij, "dail", 1.0, self.Xakic, "akic")
        C(1.0, Vabij, "dclk", Tai, "di", 0.0, Y, "clki")                       # :208
        C(-1.0, Y, "clki", Tai, "al", 1.0, self.Xakic, "akic")
        C(1.0, Vabij, "dclk", Tabij, "adil", 1.0, self.Xakic, "akic")
        C(-0.5, Vabij, "cdlk", Tabij, "adil", 1.0, self.Xakic, "akic")
        # Xakci (:213-217)
        A(1.0, Vaibj, "akci", 0.0, self.Xakci, "akci")
        C(-1.0, Vijka, "klic", Tai, "al", 1.0, self.Xakci, "akci")
        C(1.0, Vabci, "adck", Tai, "di", 1.0, self.Xakci, "akci")
        C(-0.5, Vabij, "cdlk", Tabij, "dail", 1.0, self.Xakci, "akci")
        C(1.0, Vabij, "cdlk", Tai, "di", 0.0, Y, "clki")                       # :217
        C(-1.0, Y, "clki", Tai, "al", 1.0, self.Xakci, "akci")
        # :220-224
        C(2.0, self.Xakic, "akic", Tabij, "cbkj", 1.0, Rabij, "abij")
        C(-1.0, self.Xakic, "akic", Tabij, "bckj", 1.0, Rabij, "abij")
        C(-1.0, self.Xakci, "akci", Tabij, "cbkj", 1.0, Rabij, "abij")
        C(-1.0, self.Xakci, "bkci", Tabij, "ackj", 1.0, Rabij, "abij")
        # permutation operator (:228-229)
        A(1.0, Rabij, "abij", 0.0, self.S, "abij")
        A(1.0, self.S, "baji", 1.0, Rabij, "abij")
        A(1.0, Vabij, "abij", 1.0, Rabij, "abij")                              # :238
        # Xklij (:241-245)
        A(1.0, Vijkl, "klij", 0.0, self.Xklij, "klij")
        C(1.0, Vijka, "klic", Tai, "cj", 1.0, self.Xklij, "klij")
        C(1.0, Vijka, "lkjc", Tai, "ci", 1.0, self.Xklij, "klij")
        C(1.0, Vabij, "cdkl", X, "cdij", 1.0, self.Xklij, "klij")
        C(1.0, self.Xklij, "klij", X, "abkl", 1.0, Rabij, "abij")              # :248-251
        # Xabcd (:254-256)
        A(1.0, Vabcd, "abcd", 0.0, self.Xabcd, "abcd")
        C(-1.0, Vabci, "cdak", Tai, "bk", 1.0, self.Xabcd, "abcd")
        C(-1.0, Vabci, "dcbk", Tai, "ak", 1.0, self.Xabcd, "abcd")
        C(1.0, self.Xabcd, "abcd", X, "cdij", 1.0, Rabij, "abij")              # :259-260
        # T1 equations (:270-293)
        C(1.0, self.Kac, "ac", Tai, "ci", 0.0, Rai, "ai")
        C(-1.0, self.Kki, "ki", Tai, "ak", 1.0, Rai, "ai")
        C(2.0, Vabij, "cdkl", Tai, "dl", 0.0, self.Kck, "ck")
        C(-1.0, Vabij, "cdlk", Tai, "dl", 1.0, self.Kck, "ck")
        C(2.0, self.Kck, "ck", Tabij, "caki", 1.0, Rai, "ai")
        C(-1.0, self.Kck, "ck", Tabij, "caik", 1.0, Rai, "ai")
        C(1.0, self.Kck, "ck", Tai, "ci", 0.0, self.Zki, "ki")                 # :280
        C(1.0, self.Zki, "ki", Tai, "ak", 1.0, Rai, "ai")
        C(2.0, Vabij, "acik", Tai, "ck", 1.0, Rai, "ai")
        C(-1.0, Vaibj, "akci", Tai, "ck", 1.0, Rai, "ai")
        C(2.0, Vabci, "cdak", Tabij, "cdik", 1.0, Rai, "ai")
        C(-1.0, Vabci, "dcak", Tabij, "cdik", 1.0, Rai, "ai")
        C(1.0, self.Mac, "ac", Tai, "ci", 1.0, Rai, "ai")                      # :286-287 = (Lac - Kac) Tai
        C(-2.0, Vijka, "klic", Tabij, "ackl", 1.0, Rai, "ai")
        C(1.0, Vijka, "lkic", Tabij, "ackl", 1.0, Rai, "ai")
        C(-1.0, self.Mki, "ki", Tai, "ak", 1.0, Rai, "ai")                     # :290-291 = -(Lki - Kki) Tai

    def energy(self, Tai, Tabij):
        """getEnergy (:160-178), spins = 2: direct 2 X.V minus exchange X.V["baij"]."""
        self.build_x(Tai, Tabij)
        dire = 2.0 * self.eng.dot(self.X, self.V["PPHH"])
        exce = -1.0 * self.eng.dot(self.X, self.Vx)
        return dire + exce, dire, exce

    # ------------------------------------------------------------------ the solver loop
    def solve(self, mixer="LinearMixer", max_residua=4, mixing_ratio=1.0, max_iterations=DEFAULT_MAX_ITERATIONS,
              energy_convergence=DEFAULT_ENERGY_CONVERGENCE, amplitudes_convergence=DEFAULT_AMPLITUDES_CONVERGENCE,
              level_shift=DEFAULT_LEVEL_SHIFT, log=None):
        """ClusterSinglesDoublesAlgorithm::run<double> (:37-128).  Returns dict(energy, T1, T2,
        iterations, converged); not converging is reported, not raised (the reference logs a WARNING
        and stores the amplitudes, :120-124)."""
        eng, o, v = self.eng, self.no, self.nv
        if mixer == "DiisMixer":
            mix = _DiisMixer(eng, max_residua)
        elif mixer == "LinearMixer":
            mix = _LinearMixer(eng, mixing_ratio)
        else:
            raise ValueError(f"Mixer not implemented: {mixer}")                 # :50-54
        T = [eng.tensor((v, o)), eng.tensor((v, v, o, o))]
        own_T = True
        e = prev = 0.0
        converged, it = False, -1
        for it in range(int(max_iterations)):
            R = [eng.tensor((v, o)), eng.tensor((v, v, o, o))]
            self.residuum(it, T[0], T[1], R[0], R[1])
            for r, t in zip(R, T):                                             # estimateAmplitudesFromResiduum
                eng.excitation_divide(r, t, self.epsi, self.epsa, level_shift)
            change = [eng.tensor((v, o)), eng.tensor((v, v, o, o))]            # amplitudesChange = estimate - amplitudes
            for c, r, t in zip(change, R, T):
                eng.add(1.0, r, _IDX[len(r.shape)], 0.0, c, _IDX[len(r.shape)])
                eng.add(-1.0, t, _IDX[len(r.shape)], 1.0, c, _IDX[len(r.shape)])
            dd = sum(eng.dot(c, c) for c in change)
            mix.append(R, change)          # the mixer owns the estimates and their residua from here on
            if own_T:                      # the initial (zero) amplitudes; later ones belong to the mixer
                for t in T:
                    t.free()
                own_T = False
            T = mix.get()
            e, dire, exce = self.energy(T[0], T[1])
            if log:
                log(f"iteration: {it + 1}  energy= {e:.10f}  dir= {dire:.10f}  exc= {exce:.10f}")
            tt = sum(eng.dot(t, t) for t in T)
            if abs((e - prev) / e) < abs(energy_convergence) and abs(dd / tt) < abs(amplitudes_convergence ** 2):
                converged = True
                break
            prev = e
        if int(max_iterations) == 0:
            e = self.energy(T[0], T[1])[0]
        out = dict(energy=float(e), T1=T[0].get(), T2=T[1].get(), iterations=it + 1, converged=converged,
                   stats=eng.stats())
        return out


def solve_ccsd(epsi, epsa, integrals, device: int = 0, **kw):
    """One-call form: integral blocks (host arrays) -> converged amplitudes (host arrays)."""
    with CcsdSolver(epsi, epsa, integrals, device) as s:
        return s.solve(**kw)
