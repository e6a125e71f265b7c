"""Spin-orbital perturbative triples (UPerturbativeTriples.cxx:19-305): the literal NumPy
restatement is pinned by reproducing the closed-shell (T) energy on the spin-orbital image of a closed-shell
system; the device path (tensor engine) is checked against the restatement."""
import numpy as np
import pytest

from sisi4s_b200 import synthetic as S


def _closed_shell(o=2, v=4, seed=1):
    from oracle import ccsd_ref as R
    epsi, epsa = S.eigenenergies(o, v)
    gamma = S.make_vertex(o, v, seed=4, nf=9, kappa=0.5)
    V = R.integral_blocks(gamma, o, v)
    rng = np.random.default_rng(seed)
    X = 0.1 * rng.standard_normal((v, v, o, o))
    return epsi, epsa, 0.1 * rng.standard_normal((v, o)), X + X.transpose(1, 0, 3, 2), V, gamma


def test_spin_orbital_restatement_reproduces_the_closed_shell_energy():
    """E(T) of the closed-shell loop form (oracle/pt_oracle.py, pinned by the reference's recorded UEG
    energy) == the spin-orbital full-tensor form on antisymmetrised integrals, singles included."""
    from oracle import pt_oracle as O, upt_oracle as U
    epsi, epsa, T1, T2, V, gamma = _closed_shell()
    e_cs = O.triples_loop(epsi, epsa, T1, T2, V["PPHH"], V["HHHP"], V["PPPH"])
    e_u = U.triples(*U.spin_orbital_image(epsi, epsa, T1, T2, gamma))
    assert abs(e_u - e_cs) <= 1e-14 * max(1.0, abs(e_cs)) and abs(e_cs) > 1e-3


@pytest.mark.gpu
def test_device_spin_orbital_triples_match_the_restatement():
    from oracle import upt_oracle as U
    from sisi4s_b200.plan import run_plan_file  # noqa: F401  (registers the step)
    from sisi4s_b200.triples import AlgorithmFactory
    from sisi4s_b200.triples_spin_orbital import spin_orbital_triples_energy
    epsi, epsa, T1, T2, V, gamma = _closed_shell(o=2, v=5, seed=3)
    args = U.spin_orbital_image(epsi, epsa, T1, T2, gamma)
    want = U.triples(*args)
    got = spin_orbital_triples_energy(*args)
    assert abs(got - want) <= 1e-13 * max(1.0, abs(want))
    # unsymmetric random spin-orbital inputs: the statements themselves, no symmetry assumed
    rng = np.random.default_rng(11)
    o, v = 3, 4
    ei, ea = S.eigenenergies(o, v)
    r = lambda *s: np.asfortranarray(0.3 * rng.standard_normal(s))
    raw = (ei, ea, r(v, o), r(v, v, o, o), r(v, v, o, o), r(o, o, o, v), r(v, v, v, o))
    assert abs(spin_orbital_triples_energy(*raw) - U.triples(*raw)) <= 1e-12
    keys = ("HoleEigenEnergies", "ParticleEigenEnergies", "CcsdSinglesAmplitudes", "CcsdDoublesAmplitudes",
            "PPHHCoulombIntegrals", "HHHPCoulombIntegrals", "PPPHCoulombIntegrals")
    data = dict(zip(keys, raw))
    amap = {k: "$" + k for k in keys}
    amap["PerturbativeTriplesEnergy"] = "$E"
    AlgorithmFactory.create("UPerturbativeTriples", amap, data).run()
    assert abs(data["E"] - U.triples(*raw)) <= 1e-12
