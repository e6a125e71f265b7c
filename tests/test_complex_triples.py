"""Complex closed-shell (T) (first slice): the device path -- the real step's fused
kernel driven twice with stacked real / imaginary parts (sisi4s_b200/triples_complex.py) -- against the
NumPy restatement of CcsdPerturbativeTriplesComplex.cxx:166-271 (oracle/pt_complex_oracle.py)."""
import numpy as np
import pytest

from sisi4s_b200 import synthetic as S


def _complex_inputs(o, v, nf, seed):
    rng = np.random.default_rng(seed)
    c = lambda *shape: np.asfortranarray(0.3 * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)))
    epsi, epsa = S.eigenenergies(o, v)
    return epsi, epsa, c(v, o), c(v, v, o, o), c(v, v, o, o), c(v, o, o, o), c(nf, o + v, o + v)


def test_complex_oracle_reduces_to_the_real_oracle():
    """CPU: with real inputs the complex restatement must give the real step's energies (oracle/pt_oracle.py,
    pinned by the reference's recorded UEG (T) energy); PHHH["clkj"] = HHHP["jklc"]
    (CoulombIntegralsFromVertex.cxx:539)."""
    from oracle import pt_complex_oracle as OC, pt_oracle as O
    inp = S.make_inputs(3, 5, seed=4, kind="vertex")
    e_ref, per_ref = O.triples_loop(*inp.args(), return_per_triple=True)
    greal = np.concatenate([inp.Gamma.real, inp.Gamma.imag], axis=0).astype(complex)   # same integrals, real vertex
    e, per = OC.triples_complex(inp.epsi, inp.epsa, inp.T1, inp.T2, inp.Vpphh, np.einsum("jklc->clkj", inp.Vhhhp),
                                greal, return_per_triple=True)
    assert abs(e.real - e_ref) <= 1e-14 and abs(e.imag) <= 1e-14
    assert np.abs(per.real - per_ref).max() <= 1e-14


@pytest.mark.gpu
@pytest.mark.parametrize("o,v,nf,seed", [(2, 5, 4, 1), (3, 17, 6, 2), (4, 20, 9, 3)])
def test_complex_triples_match_the_oracle(o, v, nf, seed):
    from oracle import pt_complex_oracle as OC
    from sisi4s_b200.triples_complex import complex_triples_energy
    epsi, epsa, T1, T2, P, U, g = _complex_inputs(o, v, nf, seed)
    e_ref, per_ref = OC.triples_complex(epsi, epsa, T1, T2, P, U, g, return_per_triple=True)
    e, per = complex_triples_energy(epsi, epsa, T1, T2, P, U, g, return_per_triple=True)
    scale = max(1.0, np.abs(per_ref).max())
    assert np.abs(per - per_ref.real).max() <= 1e-11 * scale, (per, per_ref.real)
    assert abs(e - e_ref.real) <= 1e-11 * max(1.0, abs(e_ref))


@pytest.mark.gpu
def test_complex_path_with_real_inputs_is_the_real_step():
    """Real amplitudes / integrals through the complex driver = the real engine on the same inputs."""
    from sisi4s_b200.triples import TriplesEngine
    from sisi4s_b200.triples_complex import complex_triples_energy
    inp = S.make_inputs(5, 19, seed=2026, kind="vertex")
    with TriplesEngine(5, 19) as eng:
        eng.set_inputs(*inp.args())
        base = eng.run()
    greal = np.concatenate([inp.Gamma.real, inp.Gamma.imag], axis=0)
    e, per = complex_triples_energy(inp.epsi, inp.epsa, inp.T1, inp.T2, inp.Vpphh,
                                       np.einsum("jklc->clkj", inp.Vhhhp), greal, return_per_triple=True)
    assert abs(e - base.energy) <= 1e-12
    assert np.abs(per - base.per_triple).max() <= 1e-12


@pytest.mark.gpu
def test_complex_plan_step():
    from oracle import pt_complex_oracle as OC
    from sisi4s_b200.plan import run_plan_file  # noqa: F401  (registers the step)
    from sisi4s_b200.triples import AlgorithmFactory, SisiException
    epsi, epsa, T1, T2, P, U, g = _complex_inputs(2, 6, 5, 9)
    data = dict(HoleEigenEnergies=epsi, ParticleEigenEnergies=epsa, CoulombVertex=g, CcsdSinglesAmplitudes=T1,
                CcsdDoublesAmplitudes=T2, PPHHCoulombIntegrals=P, PHHHCoulombIntegrals=U, CcsdEnergy=-1.25)
    args = {k: "$" + k for k in data}
    args["CcsdPerturbativeTriplesComplexEnergy"] = "$E"
    AlgorithmFactory.create("CcsdPerturbativeTriplesComplex", args, data).run()
    assert abs(data["E"] - (-1.25 + OC.triples_complex(epsi, epsa, T1, T2, P, U, g).real)) <= 1e-10
    del args["CcsdEnergy"]
    with pytest.raises(SisiException, match="Missing argument: CcsdEnergy"):
        AlgorithmFactory.create("CcsdPerturbativeTriplesComplex", args, data).run()
