"""The C++ plugin classes (sisi4s_b200/csrc/*Gpu.cxx) RUN through the reference's own argument machinery.

oracle/_ref/plugin_harness (oracle/Makefile, oracle/harness/) links the plugin classes with the reference's
Algorithm.cxx / Data.cxx / DryTensor.cxx / util/Log.cxx / util/Emitter.cxx -- compiled where they lie under
/root/reference -- and with single-process stand-ins for the three absent third-party headers (<ctf.hpp>,
<mpi.h>, <yaml-cpp/yaml.h>).  So AlgorithmFactory registration, getTensorArgument / getRealArgument /
setRealArgument on real Data objects, Tensor::read_all order, option handling and error propagation of the
drop-in are executed, not only syntax-checked; energies are compared with the oracles.  The binary is built
in the build container (where the reference tree is) and travels to the GPU box."""
import os
import re
import subprocess

import numpy as np
import pytest

from sisi4s_b200 import synthetic as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "plugin_harness")
pytestmark = pytest.mark.skipif(not os.path.isfile(HARNESS), reason="oracle/_ref/plugin_harness not built (needs the reference tree)")
ABS_TOL = 1e-9


def _run(tmp_path, algorithm, entries, dry=False):
    """entries: (kind, name, value) with kind in tensor / ctensor / real / integer / text / out / tout."""
    lines = []
    for kind, name, value in entries:
        if kind in ("tensor", "ctensor"):
            a = np.asfortranarray(value, dtype=np.float64 if kind == "tensor" else np.complex128)
            path = tmp_path / f"{name}.bin"
            a.ravel(order="F").tofile(path)
            lines.append(f"{kind} {name} {path} " + " ".join(str(n) for n in a.shape))
        elif kind == "out":
            lines.append(f"out {name}")
        elif kind == "tout":
            lines.append(f"tout {name} {tmp_path / (name + '.out')}")
        else:
            lines.append(f"{kind} {name} {value!r}" if kind == "real" else f"{kind} {name} {value}")
    plan = tmp_path / "plan.txt"
    plan.write_text("\n".join(lines) + "\n")
    res = subprocess.run([HARNESS, algorithm, str(plan)] + (["--dry"] if dry else []), capture_output=True, text=True,
                         timeout=300, cwd=str(tmp_path))
    values = {m.group(1): float(m.group(2)) for m in re.finditer(r"^(\w+) = (\S+)$", res.stdout, re.M)}
    return res, values


def _triples_entries(inp, ccsd=-0.25):
    return [("tensor", "HoleEigenEnergies", inp.epsi), ("tensor", "ParticleEigenEnergies", inp.epsa),
            ("tensor", "CcsdSinglesAmplitudes", inp.T1), ("tensor", "CcsdDoublesAmplitudes", inp.T2),
            ("tensor", "PPHHCoulombIntegrals", inp.Vpphh), ("tensor", "HHHPCoulombIntegrals", inp.Vhhhp),
            ("tensor", "PPPHCoulombIntegrals", inp.Vppph), ("real", "CcsdEnergy", ccsd)]


# ------------------------------------------------------------ host side only
def test_reference_argument_errors_reach_the_caller(tmp_path):
    """Missing mandatory CcsdEnergy: the message is the reference's own (Algorithm.cxx:37-45), thrown from the
    reference's code."""
    inp = S.make_inputs(2, 3, seed=1, kind="vertex")
    entries = [e for e in _triples_entries(inp) if e[1] != "CcsdEnergy"] + [("out", "CcsdPerturbativeTriplesEnergy", None)]
    res, values = _run(tmp_path, "CcsdPerturbativeTriplesGpu", entries)
    out = res.stdout
    if "no CUDA device" in out:                      # the device check comes first on a box without a GPU
        assert res.returncode == 1
    else:
        assert res.returncode == 1 and "Missing argument: CcsdEnergy" in out and "Algorithm.cxx" in out
    assert not values
    res, _ = _run(tmp_path, "NoSuchAlgorithm", entries)
    assert res.returncode == 1 and "unknown algorithm" in res.stdout


def test_dry_run_reports_the_library_estimate(tmp_path):
    """dryRun() on DryTensor arguments (no GPU needed): the logged figure is pt_estimate_device_bytes."""
    from sisi4s_b200 import _lib
    o, v = 40, 300
    z = lambda *s: np.zeros(s)
    entries = [("tensor", "HoleEigenEnergies", z(o)), ("tensor", "ParticleEigenEnergies", z(v)), ("integer", "holeBlock", 8)]
    res, _ = _run(tmp_path, "CcsdPerturbativeTriplesGpu", entries, dry=True)
    assert res.returncode == 0, res.stdout + res.stderr
    m = re.search(r"device memory per GPU=([0-9.e+-]+) GB", res.stdout)
    want = _lib.load().pt_estimate_device_bytes(o, v, 0, 8) / 1e9
    assert m and abs(float(m.group(1)) - want) <= 1e-4 * want


def test_no_cpu_fallback_in_the_plugin_class(tmp_path):
    from conftest import gpu_available
    if gpu_available():
        pytest.skip("CUDA device present")
    inp = S.make_inputs(2, 3, seed=1, kind="vertex")
    res, values = _run(tmp_path, "CcsdPerturbativeTriplesGpu", _triples_entries(inp) + [("out", "CcsdPerturbativeTriplesEnergy", None)])
    assert res.returncode == 1 and "no CUDA device (there is no CPU fallback)" in res.stdout and not values


# ------------------------------------------------------------ on the GPU
@pytest.mark.gpu
def test_triples_plugin_class_ppph_contract(tmp_path):
    """PerturbativeTriples contract (PPPH tensor given), both output spellings, slab-wise Tensor::slice upload."""
    from oracle import pt_oracle as O
    inp = S.make_inputs(5, 19, seed=7, kind="vertex")
    e_ref = O.triples_loop(*inp.args())
    res, values = _run(tmp_path, "CcsdPerturbativeTriplesGpu", _triples_entries(inp, ccsd=-0.5) + [
        ("out", "CcsdPerturbativeTriplesEnergy", None), ("out", "PerturbativeTriplesEnergy", None)])
    assert res.returncode == 0, res.stdout + res.stderr
    assert abs(values["CcsdPerturbativeTriplesEnergy"] - (-0.5 + e_ref)) <= ABS_TOL
    assert values["PerturbativeTriplesEnergy"] == values["CcsdPerturbativeTriplesEnergy"]
    # a wrong shape is refused with the argument's name
    bad = [(k, n, (v_.T if n == "CcsdSinglesAmplitudes" else v_)) for k, n, v_ in _triples_entries(inp)]
    res, values = _run(tmp_path, "CcsdPerturbativeTriplesGpu", bad + [("out", "CcsdPerturbativeTriplesEnergy", None)])
    assert res.returncode == 1 and "Incompatible shape of argument: CcsdSinglesAmplitudes" in res.stdout and not values


@pytest.mark.gpu
@pytest.mark.parametrize("options", [[], [("integer", "integralsFromVertex", 1), ("integer", "holeBlock", 2)]])
def test_triples_plugin_class_vertex_contract(tmp_path, options):
    """CcsdPerturbativeTriples contract (CoulombVertex, complex Tensor), resident and hole-blocked with all
    integrals from the vertex."""
    from oracle import pt_oracle as O
    inp = S.make_inputs(5, 19, seed=8, kind="vertex")
    e_ref = O.triples_loop(*inp.args())
    entries = [e for e in _triples_entries(inp, ccsd=0.125) if e[1] != "PPPHCoulombIntegrals"]
    if options:
        entries = [e for e in entries if e[1] not in ("PPHHCoulombIntegrals", "HHHPCoulombIntegrals")]
    entries += [("ctensor", "CoulombVertex", inp.Gamma)] + options + [("out", "CcsdPerturbativeTriplesEnergy", None)]
    res, values = _run(tmp_path, "CcsdPerturbativeTriplesGpu", entries)
    assert res.returncode == 0, res.stdout + res.stderr
    assert abs(values["CcsdPerturbativeTriplesEnergy"] - (0.125 + e_ref)) <= ABS_TOL


@pytest.mark.gpu
def test_complex_triples_plugin_class(tmp_path):
    from oracle import pt_complex_oracle as OC
    o, v, nf = 3, 17, 6
    rng = np.random.default_rng(2)
    c = lambda *shape: np.asfortranarray(0.3 * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)))
    epsi, epsa = S.eigenenergies(o, v)
    T1, T2, P, U, g = c(v, o), c(v, v, o, o), c(v, v, o, o), c(v, o, o, o), c(nf, o + v, o + v)
    e_ref = OC.triples_complex(epsi, epsa, T1, T2, P, U, g)
    res, values = _run(tmp_path, "CcsdPerturbativeTriplesComplexGpu", [
        ("tensor", "HoleEigenEnergies", epsi), ("tensor", "ParticleEigenEnergies", epsa),
        ("ctensor", "CcsdSinglesAmplitudes", T1), ("ctensor", "CcsdDoublesAmplitudes", T2),
        ("ctensor", "PPHHCoulombIntegrals", P), ("ctensor", "PHHHCoulombIntegrals", U), ("ctensor", "CoulombVertex", g),
        ("real", "CcsdEnergy", 1.0), ("out", "CcsdPerturbativeTriplesComplexEnergy", None)])
    assert res.returncode == 0, res.stdout + res.stderr
    want = 1.0 + complex(e_ref).real
    assert abs(values["CcsdPerturbativeTriplesComplexEnergy"] - want) <= 1e-11 * max(1.0, abs(want))


@pytest.mark.gpu
def test_spin_orbital_triples_plugin_class(tmp_path):
    from oracle import upt_oracle as U
    rng = np.random.default_rng(11)
    o, v = 3, 4
    ei, ea = S.eigenenergies(o, v)
    r = lambda *s: np.asfortranarray(0.3 * rng.standard_normal(s))
    raw = (ei, ea, r(v, o), r(v, v, o, o), r(v, v, o, o), r(o, o, o, v), r(v, v, v, o))
    keys = ("HoleEigenEnergies", "ParticleEigenEnergies", "CcsdSinglesAmplitudes", "CcsdDoublesAmplitudes",
            "PPHHCoulombIntegrals", "HHHPCoulombIntegrals", "PPPHCoulombIntegrals")
    res, values = _run(tmp_path, "UPerturbativeTriplesGpu", [("tensor", k, a) for k, a in zip(keys, raw)]
                       + [("out", "PerturbativeTriplesEnergy", None)])
    assert res.returncode == 0, res.stdout + res.stderr
    assert abs(values["PerturbativeTriplesEnergy"] - U.triples(*raw)) <= 1e-12


@pytest.mark.gpu
def test_ccsd_plugin_class(tmp_path):
    """CcsdEnergyFromCoulombIntegralsGpu from the vertex alone: energy and stored amplitudes (allocatedTensorArgument)
    against the literal NumPy restatement of the reference's solver."""
    from oracle import ccsd_ref as R
    o, v = 3, 6
    epsi, epsa = S.eigenenergies(o, v)
    gamma = S.make_vertex(o, v, seed=4, nf=14, kappa=0.55)
    V = R.integral_blocks(gamma, o, v)
    kw = dict(mixer="DiisMixer", max_iterations=60, energy_convergence=1e-11, amplitudes_convergence=1e-10)
    want = R.solve(epsi, epsa, V, **kw)
    res, values = _run(tmp_path, "CcsdEnergyFromCoulombIntegralsGpu", [
        ("tensor", "HoleEigenEnergies", epsi), ("tensor", "ParticleEigenEnergies", epsa), ("ctensor", "CoulombVertex", gamma),
        ("text", "mixer", "DiisMixer"), ("integer", "maxIterations", 60), ("real", "energyConvergence", 1e-11),
        ("real", "amplitudesConvergence", 1e-10), ("out", "CcsdEnergy", None),
        ("tout", "CcsdSinglesAmplitudes", None), ("tout", "CcsdDoublesAmplitudes", None)])
    assert res.returncode == 0, res.stdout + res.stderr
    assert abs(values["CcsdEnergy"] - want["energy"]) <= 1e-10
    t1 = np.fromfile(tmp_path / "CcsdSinglesAmplitudes.out").reshape((v, o), order="F")
    t2 = np.fromfile(tmp_path / "CcsdDoublesAmplitudes.out").reshape((v, v, o, o), order="F")
    assert np.abs(t1 - want["T1"]).max() <= 1e-9 and np.abs(t2 - want["T2"]).max() <= 1e-9
