"""Pins oracle/ccsd_ref.py -- the literal NumPy restatement of the reference's closed-shell CCSD
residuum, mixers and convergence test -- against the CCSD energy the reference records for its UEG
test system (integration-tests/tests/cc4s/ueg/rs1.0-7occ-26virt/cc4s.correct.out.yaml:153) and against
the independent textbook spin-orbital solver (oracle/ccsd.py)."""
import os

import numpy as np
import pytest

from oracle import ccsd_ref as R
from sisi4s_b200 import ueg

REF_CCSD = -0.39269658954585018
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ueg_rs1_no7_nv26.npz")


@pytest.fixture(scope="module")
def ueg_system():
    epsi, epsa, gamma = ueg.make_ueg(7, 26, 1.0)
    return epsi, epsa, R.integral_blocks(gamma, 7, 26)


def test_reference_residuum_reproduces_recorded_ccsd_energy(ueg_system):
    epsi, epsa, V = ueg_system
    res = R.solve(epsi, epsa, V, mixer="DiisMixer", max_residua=4, max_iterations=50,
                  energy_convergence=1e-10, amplitudes_convergence=1e-10)
    assert res["converged"]
    assert abs(res["energy"] - REF_CCSD) < 1e-8            # the reference converged to 1e-8
    amps = np.load(GOLD)                                   # converged by the spin-orbital solver
    assert np.abs(res["T2"] - amps["T2"]).max() < 1e-8
    assert np.abs(res["T1"]).max() < 1e-12                 # momentum conservation: no singles in the UEG


def test_linear_mixer_and_reference_defaults(ueg_system):
    """Reference defaults: LinearMixer, 16 iterations, 1e-6 / 1e-5 relative criteria
    (ClusterSinglesDoublesAlgorithm.hpp); non-convergence is not an error there."""
    epsi, epsa, V = ueg_system
    res = R.solve(epsi, epsa, V)
    assert res["iterations"] <= 16
    assert abs(res["energy"] - REF_CCSD) < 1e-4


def test_first_iteration_is_mp2(ueg_system):
    epsi, epsa, V = ueg_system
    res = R.solve(epsi, epsa, V, max_iterations=1)
    assert abs(res["energy"] - (-0.36143028565969504)) < 1e-13   # cc4s.correct.out.yaml:128


def test_agrees_with_spin_orbital_solver_when_singles_do_not_vanish():
    """A system without momentum conservation (T1 != 0): the reference's closed-shell equations and the
    textbook spin-orbital equations (oracle/ccsd.py) converge to the same amplitudes and energy."""
    from oracle import ccsd as SO
    from sisi4s_b200 import synthetic as S
    o, v = 3, 6
    epsi, epsa = S.eigenenergies(o, v)
    gamma = S.make_vertex(o, v, seed=4, nf=14, kappa=0.55)
    V = R.integral_blocks(gamma, o, v)
    a = R.solve(epsi, epsa, V, mixer="DiisMixer", max_iterations=80, energy_convergence=1e-12,
                amplitudes_convergence=1e-11)
    b = SO.solve(epsi, epsa, gamma, tol=1e-12)
    assert a["converged"]
    assert np.abs(a["T1"]).max() > 1e-4
    assert abs(a["energy"] - b["energy"]) < 1e-10
    assert np.abs(a["T1"] - b["T1"]).max() < 1e-8 and np.abs(a["T2"] - b["T2"]).max() < 1e-8
