"""Upstream steps of the (T) path as plan steps (first slices): integral blocks
from the vertex (CoulombIntegralsFromVertex.cxx index strings) and the CCSD solver, checked on the
reference's UEG test system against its recorded energies (cc4s.correct.out.yaml:128,153)."""
import numpy as np
import pytest

from sisi4s_b200 import synthetic as S
from sisi4s_b200.plan import run_plan_file
from sisi4s_b200.triples import AlgorithmFactory, SisiException

REF_MP2 = -0.36143028565969504
REF_CCSD = -0.39269658954585018


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


def test_ueg_plan_reproduces_reference_mp2_and_ccsd(tmp_path, monkeypatch):
    """UegVertexGenerator -> CoulombIntegralsFromVertex -> CcsdEnergyFromCoulombIntegralsReference as a
    YAML plan (CPU device asked for explicitly): MP2 and CCSD energies of cc4s.correct.out.yaml."""
    monkeypatch.chdir(tmp_path)
    open("in.yaml", "w").write("""
- name: UegVertexGenerator
  in: {No: 7, Nv: 26, rs: 1.0}
  out: {CoulombVertex: $CoulombVertex, HoleEigenEnergies: $HoleEigenEnergies, ParticleEigenEnergies: $ParticleEigenEnergies}
- name: CoulombIntegralsFromVertex
  in: {CoulombVertex: $CoulombVertex, HoleEigenEnergies: $HoleEigenEnergies, ParticleEigenEnergies: $ParticleEigenEnergies, complex: 0}
  out: {PPHHCoulombIntegrals: $PPHHCoulombIntegrals, HHHPCoulombIntegrals: $HHHPCoulombIntegrals}
- name: CcsdEnergyFromCoulombIntegralsReference
  in:
    mixer: DiisMixer
    maxResidua: 4
    maxIterations: 50
    energyConvergence: 1e-8
    amplitudesConvergence: 1e-8
    device: cpu
    CoulombVertex: $CoulombVertex
    HoleEigenEnergies: $HoleEigenEnergies
    ParticleEigenEnergies: $ParticleEigenEnergies
    PPHHCoulombIntegrals: $PPHHCoulombIntegrals
  out: {CcsdEnergy: $CcsdEnergy, CcsdSinglesAmplitudes: $CcsdSinglesAmplitudes, CcsdDoublesAmplitudes: $CcsdDoublesAmplitudes}
""")
    data = run_plan_file("in.yaml", log=lambda *_: None)
    from sisi4s_b200 import ueg
    assert abs(ueg.mp2_energy(data["HoleEigenEnergies"], data["ParticleEigenEnergies"], data["PPHHCoulombIntegrals"]) - REF_MP2) < 1e-13
    assert abs(data["CcsdEnergy"] - REF_CCSD) < 1e-8
    assert data["CcsdDoublesAmplitudes"].shape == (26, 26, 7, 7) and np.abs(data["CcsdSinglesAmplitudes"]).max() < 1e-12
