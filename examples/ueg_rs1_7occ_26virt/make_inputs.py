"""Writes the input files of the UEG example (rs = 1.0, 7 occupied / 26 virtual states) into the current
directory, in the file formats of the reference, with the file names its own test plan uses
(integration-tests/tests/cc4s/ueg/rs1.0-7occ-26virt/in.yaml: EigenEnergies.yaml, CoulombVertex.yaml +
.elements), plus what the reference's CCSD step would hand to the (T) step (amplitudes and the two
integral blocks, as TENS binaries a `TensorWriter mode: binary` step of sisi4s produces).

    python examples/ueg_rs1_7occ_26virt/make_inputs.py && python -m sisi4s_b200 in.yaml

The Hamiltonian comes from sisi4s_b200/ueg.py (restating UegVertexGenerator.cxx), the amplitudes from the
committed fixture tests/golden/ueg_rs1_no7_nv26.npz (oracle/ccsd.py).  Expected output:
CcsdPerturbativeTriplesEnergy = E(CCSD) + E(T) = -0.39269658979 - 0.00630196257, the values recorded in
the reference's cc4s.correct.out.yaml:153,169.
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from sisi4s_b200 import synthetic as S  # noqa: E402
from sisi4s_b200 import ueg  # noqa: E402
from sisi4s_b200 import tensor_io as TIO  # noqa: E402

no, nv = 7, 26
epsi, epsa, gamma = ueg.make_ueg(no, nv, 1.0)
vpphh, vhhhp, _ = S.integrals_from_vertex(gamma, no, nv, with_ppph=False)
amps = np.load(os.path.join(ROOT, "tests", "golden", "ueg_rs1_no7_nv26.npz"))

energies = ", ".join(repr(float(x)) for x in np.concatenate([epsi, epsa]))
fermi = 0.5 * (float(epsi.max()) + float(epsa.min()))
with open("EigenEnergies.yaml", "w") as f:
    f.write(f"version: 100\ntype: Tensor\nscalarType: Real64\nmetaData:\n  fermiEnergy: {fermi!r}\n"
            f"  energies: [{energies}]\n")
TIO.write_cc4s("CoulombVertex.yaml", gamma, binary=True, axis_types=["AuxiliaryField", "State", "State"])
TIO.write_binary("CcsdSinglesAmplitudes.bin", amps["T1"])
TIO.write_binary("CcsdDoublesAmplitudes.bin", amps["T2"])
TIO.write_binary("PPHHCoulombIntegrals.bin", vpphh)
TIO.write_binary("HHHPCoulombIntegrals.bin", vhhhp)
if os.path.abspath(os.getcwd()) != HERE:
    shutil.copy(os.path.join(HERE, "in.yaml"), "in.yaml")
print("E(CCSD) of the stored amplitudes:", float(amps["ccsd_energy"]))
