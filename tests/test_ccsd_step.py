"""Upstream steps of the (T) path on the device: the tensor-contraction engine,
every integral block of CoulombIntegralsFromVertex, the closed-shell CCSD residuum / solver of the
reference -- against NumPy, against the literal oracle (oracle/ccsd_ref.py) and against the energies the
reference records for its UEG test system (cc4s.correct.out.yaml:128,153)."""
import numpy as np
import pytest

from sisi4s_b200 import synthetic as S
from sisi4s_b200.plan import run_plan_file
from sisi4s_b200.triples import AlgorithmFactory, SisiException

pytestmark = pytest.mark.gpu
REF_MP2 = -0.36143028565969504
REF_CCSD = -0.39269658954585018


# ------------------------------------------------------------ the tensor engine
def test_contractions_match_einsum():
    """CTF-style statements C[ic] = a A[ia] B[ib] + b C[ic]: direct GEMM output, swapped operands,
    scratch + permuted add, outer products, full contractions, odd extents."""
    from sisi4s_b200.tensor_engine import DeviceTensors, TnError
    rng = np.random.default_rng(7)
    ext = dict(a=5, b=5, c=5, d=5, i=3, j=3, k=3, l=3, G=11)
    cases = [("acik", "cbkj", "abij"), ("cdkl", "adkl", "ac"), ("klij", "abkl", "abij"), ("ai", "bj", "abij"),
             ("abcd", "cdij", "abij"), ("ki", "abkj", "abij"), ("Gac", "Gbd", "abcd"), ("Gik", "Gaj", "ijka"),
             ("ck", "caki", "ai"), ("bkij", "ak", "abij"), ("cdkl", "cdkl", ""), ("ac", "ci", "ai")]
    with DeviceTensors() as eng:
        for ia, ib, ic in cases:
            A = np.asfortranarray(rng.standard_normal([ext[c] for c in ia]))
            B = np.asfortranarray(rng.standard_normal([ext[c] for c in ib]))
            C0 = np.asfortranarray(rng.standard_normal([ext[c] for c in ic])) if ic else np.array(0.37)
            tA, tB, tC = eng.tensor(A.shape, A), eng.tensor(B.shape, B), eng.tensor(C0.shape, C0)
            eng.contract(-1.5, tA, ia, tB, ib, 0.5, tC, ic)
            want = -1.5 * np.einsum(f"{ia},{ib}->{ic}", A, B) + 0.5 * C0
            assert np.abs(tC.get() - want).max() <= 1e-13 * max(1.0, np.abs(want).max()), (ia, ib, ic)
            eng.contract(2.0, tA, ia, tB, ib, 0.0, tC, ic)
            assert np.abs(tC.get() - 2.0 * np.einsum(f"{ia},{ib}->{ic}", A, B)).max() <= 1e-12, (ia, ib, ic)
            for t in (tA, tB, tC):
                t.free()
        # permuted add, dot, excitation energies
        A = np.asfortranarray(rng.standard_normal((5, 5, 3, 3)))
        C0 = np.asfortranarray(rng.standard_normal((5, 3, 5, 3)))
        tA, tC = eng.tensor(A.shape, A), eng.tensor(C0.shape, C0)
        eng.add(2.0, tA, "abij", -1.0, tC, "aibj")
        assert np.abs(tC.get() - (2.0 * A.transpose(0, 2, 1, 3) - C0)).max() <= 1e-14
        assert abs(eng.dot(tA, tA) - float(np.vdot(A, A))) <= 1e-12
        ei, ea = np.linspace(-2, -1, 3), np.linspace(1, 3, 5)
        tR, tT = eng.tensor(A.shape, A), eng.tensor(A.shape, 0.1 * A)
        eng.excitation_divide(tR, tT, eng.tensor((3,), ei), eng.tensor((5,), ea), 0.25)
        D = ea[:, None, None, None] + ea[None, :, None, None] - ei[None, None, :, None] - ei[None, None, None, :]
        assert np.abs(tR.get() - (-(A - 0.25 * 0.1 * A) / (D + 0.25))).max() <= 1e-14
        with pytest.raises(TnError, match="appears in both operands and the result"):
            eng.contract(1.0, tA, "abij", tR, "abij", 0.0, tT, "abij")


def test_integral_blocks_follow_the_reference_index_strings():
    inp = S.make_inputs(3, 5, seed=4, kind="vertex")
    g = inp.Gamma
    names = ["PPHH", "HHHH", "HHHP", "PPPP", "PPPH", "PHPH", "HPPH", "HPHP", "HPPP", "HHPH", "HHPP", "PPHP",
             "PHHP", "PHPP", "PHHH", "HPHH"]
    data = dict(CoulombVertex=g, HoleEigenEnergies=inp.epsi, ParticleEigenEnergies=inp.epsa)
    args = {k: "$" + k for k in data}
    args.update({n + "CoulombIntegrals": "$" + n for n in names})
    AlgorithmFactory.create("CoulombIntegralsFromVertex", args, data).run()
    assert np.allclose(data["PPHH"], inp.Vpphh, rtol=1e-13, atol=1e-15)
    assert np.allclose(data["HHHP"], inp.Vhhhp, rtol=1e-13, atol=1e-15)
    assert np.allclose(data["PPPH"], inp.Vppph, rtol=1e-13, atol=1e-15)
    # all blocks are sub-blocks of one <pq|rs> = G[p,r].G[q,s] for a (p,q)-symmetric vertex
    o, v = 3, 5
    V = np.einsum("Gpr,Gqs->pqrs", g.real, g.real) + np.einsum("Gpr,Gqs->pqrs", g.imag, g.imag)
    rng = {"H": slice(0, o), "P": slice(o, o + v)}
    for n in names:
        want = V[rng[n[0]], rng[n[1]], rng[n[2]], rng[n[3]]]
        assert data[n].shape == want.shape and np.allclose(data[n], want, rtol=1e-12, atol=1e-14), n
    with pytest.raises(SisiException, match="only real"):
        AlgorithmFactory.create("CoulombIntegralsFromVertex", dict(args, complex=1), data).run()


# ------------------------------------------------------------ CCSD
def _system(o=3, v=6, nf=14, kappa=0.55, seed=4):
    from oracle import ccsd_ref as R
    epsi, epsa = S.eigenenergies(o, v)
    gamma = S.make_vertex(o, v, seed=seed, nf=nf, kappa=kappa)
    return epsi, epsa, R.integral_blocks(gamma, o, v)


def test_residuum_matches_the_literal_oracle():
    """One getResiduum evaluation on random (unsymmetric, T1 != 0) amplitudes: device statements vs
    the NumPy restatement of CcsdEnergyFromCoulombIntegralsReference.cxx:29-295."""
    from oracle import ccsd_ref as R
    from sisi4s_b200.ccsd import CcsdSolver
    epsi, epsa, V = _system()
    o, v = len(epsi), len(epsa)
    rng = np.random.default_rng(3)
    Tai = np.asfortranarray(0.1 * rng.standard_normal((v, o)))
    Tabij = np.asfortranarray(0.1 * rng.standard_normal((v, v, o, o)))
    want_ai, want_abij = R.residuum(1, Tai, Tabij, V)
    with CcsdSolver(epsi, epsa, V) as s:
        r1, r2 = s.residuum(0)                            # no amplitudes given: the MP2 branch (:52-57)
        assert np.abs(r2 - V["PPHH"]).max() == 0.0 and np.abs(r1).max() == 0.0
        s.set_amplitudes(Tai, Tabij)
        r1, r2 = s.residuum(1)
        assert np.abs(r1 - want_ai).max() <= 1e-12
        assert np.abs(r2 - want_abij).max() <= 1e-12
        res = s.solve(max_iterations=0)                   # "computing energy from given amplitudes" (:116-119)
        assert abs(res["energy"] - R.energy(Tai, Tabij, V["PPHH"])) <= 1e-12
        assert np.array_equal(res["T2"], Tabij)


def test_integrals_from_the_vertex_inside_the_solver():
    """ccsd_set_vertex: the six blocks getResiduum reads, built on the device with the index strings of
    CoulombIntegralsFromVertex.cxx:395-431, equal the NumPy restatement; solving from the vertex alone gives
    the same energy as solving from the blocks."""
    from oracle import ccsd_ref as R
    from sisi4s_b200.ccsd import CcsdSolver, BLOCKS
    o, v = 3, 6
    epsi, epsa = S.eigenenergies(o, v)
    gamma = S.make_vertex(o, v, seed=4, nf=14, kappa=0.4)
    V = R.integral_blocks(gamma, o, v)
    kw = dict(mixer="DiisMixer", max_iterations=40, energy_convergence=1e-11, amplitudes_convergence=1e-10)
    with CcsdSolver(epsi, epsa, vertex=gamma) as s:
        for b in BLOCKS:
            assert np.abs(s.get_integrals(b) - V[b]).max() <= 1e-13, b
        a = s.solve(**kw)
    with CcsdSolver(epsi, epsa, V) as s:
        b = s.solve(**kw)
    assert a["converged"] and abs(a["energy"] - b["energy"]) <= 1e-12


@pytest.mark.parametrize("mixer", ["DiisMixer", "LinearMixer"])
def test_solver_matches_the_oracle_iteration_by_iteration(mixer):
    from oracle import ccsd_ref as R
    from sisi4s_b200.ccsd import solve_ccsd
    # plain linear mixing (the reference's default) only converges for the weaker coupling
    epsi, epsa, V = _system(kappa=0.55 if mixer == "DiisMixer" else 0.4)
    kw = dict(mixer=mixer, max_iterations=60, energy_convergence=1e-11, amplitudes_convergence=1e-10)
    a = R.solve(epsi, epsa, V, **kw)
    b = solve_ccsd(epsi, epsa, V, **kw)
    assert b["converged"] and a["iterations"] == b["iterations"]
    assert abs(a["energy"] - b["energy"]) <= 1e-11
    assert np.abs(a["T1"] - b["T1"]).max() <= 1e-10 and np.abs(a["T2"] - b["T2"]).max() <= 1e-10
    assert np.abs(b["T1"]).max() > 1e-5
    assert b["stats"]["launches"] > 0 and b["stats"]["flops"] > 0


def test_ueg_plan_reproduces_reference_mp2_and_ccsd(tmp_path, monkeypatch):
    """UegVertexGenerator -> CoulombIntegralsFromVertex -> CcsdEnergyFromCoulombIntegralsReference as a
    YAML plan with the reference's argument names, everything on the device: MP2 and CCSD energies of
    cc4s.correct.out.yaml."""
    monkeypatch.chdir(tmp_path)
    blocks = ["PPHH", "PHPH", "HHHH", "HHHP", "PPPH", "PPPP"]
    outs = ", ".join(f"{b}CoulombIntegrals: ${b}CoulombIntegrals" for b in blocks)
    open("in.yaml", "w").write(f"""
- name: UegVertexGenerator
  in: {{No: 7, Nv: 26, rs: 1.0}}
  out: {{CoulombVertex: $CoulombVertex, HoleEigenEnergies: $HoleEigenEnergies, ParticleEigenEnergies: $ParticleEigenEnergies}}
- name: CoulombIntegralsFromVertex
  in: {{CoulombVertex: $CoulombVertex, HoleEigenEnergies: $HoleEigenEnergies, ParticleEigenEnergies: $ParticleEigenEnergies, complex: 0}}
  out: {{{outs}}}
- name: CcsdEnergyFromCoulombIntegralsReference
  in:
    mixer: DiisMixer
    maxResidua: 4
    maxIterations: 50
    energyConvergence: 1e-8
    amplitudesConvergence: 1e-8
    HoleEigenEnergies: $HoleEigenEnergies
    ParticleEigenEnergies: $ParticleEigenEnergies
    {outs.replace(", ", chr(10) + "    ")}
  out: {{CcsdEnergy: $CcsdEnergy, CcsdSinglesAmplitudes: $CcsdSinglesAmplitudes, CcsdDoublesAmplitudes: $CcsdDoublesAmplitudes}}
""")
    data = run_plan_file("in.yaml", log=lambda *_: None)
    from sisi4s_b200 import ueg
    assert abs(ueg.mp2_energy(data["HoleEigenEnergies"], data["ParticleEigenEnergies"], data["PPHHCoulombIntegrals"]) - REF_MP2) < 1e-13
    assert abs(data["CcsdEnergy"] - REF_CCSD) < 1e-8
    assert data["CcsdDoublesAmplitudes"].shape == (26, 26, 7, 7) and np.abs(data["CcsdSinglesAmplitudes"]).max() < 1e-12


def test_reference_defaults_do_not_abort_an_unconverged_plan():
    """Reference behaviour (ClusterSinglesDoublesAlgorithm.cxx:103-124): relative criteria, LinearMixer
    by default, and a run that hits maxIterations still stores energy and amplitudes (WARNING only)."""
    epsi, epsa, V = _system()
    data = dict(HoleEigenEnergies=epsi, ParticleEigenEnergies=epsa, **{b + "CoulombIntegrals": V[b] for b in V})
    args = {k: "$" + k for k in data}
    args.update(maxIterations=2, CcsdEnergy="$CcsdEnergy", CcsdDoublesAmplitudes="$T2")
    alg = AlgorithmFactory.create("CcsdEnergyFromCoulombIntegralsReference", args, data)
    alg.run()
    assert not alg.converged and "WARNING" in alg.note
    assert np.isfinite(data["CcsdEnergy"]) and data["T2"].shape == V["PPHH"].shape
    with pytest.raises(SisiException, match="Mixer not implemented"):
        AlgorithmFactory.create("CcsdEnergyFromCoulombIntegralsReference", dict(args, mixer="NoSuchMixer"), data).run()
    with pytest.raises(SisiException, match="Missing argument: PPPPCoulombIntegrals"):
        AlgorithmFactory.create("CcsdEnergyFromCoulombIntegralsReference",
                                {k: a for k, a in args.items() if k != "PPPPCoulombIntegrals"}, data).run()
