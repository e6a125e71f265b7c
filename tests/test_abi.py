"""CPU checks of the drop-in boundary: the shared library loads and exports every
symbol include/sisi4s_pt.h declares, host-only entry points behave, and without a
CUDA device the product path fails loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest

from sisi4s_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()


def test_header_symbols_all_exported():
    _ensure_built()
    with open(os.path.join(ROOT, "include", "sisi4s_pt.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    declared = set(re.findall(r"\b(pt_[a-z0-9_]+)\s*\(", text))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_tensor_engine_header_symbols_all_exported():
    """include/sisi4s_tn.h (device tensor-contraction engine inside the same shared library)."""
    _ensure_built()
    from sisi4s_b200 import tensor_engine as TE
    with open(os.path.join(ROOT, "include", "sisi4s_tn.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    declared = set(re.findall(r"\b(tn_[a-z0-9_]+)\s*\(", text))
    assert declared == set(TE.TN_SYMBOLS), declared ^ set(TE.TN_SYMBOLS)
    lib = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().pt_estimate_device_bytes(40, 300, 0, 0) > 12e9      # host-only, no GPU needed
    # include/sisi4s_ccsd.h (device CCSD solver)
    from sisi4s_b200 import ccsd as CC
    with open(os.path.join(ROOT, "include", "sisi4s_ccsd.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    declared = set(re.findall(r"\b(ccsd_[a-z0-9_]+)\s*\(", text))
    assert declared == set(CC.CCSD_SYMBOLS), declared ^ set(CC.CCSD_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    opt = CC.CcsdOptions()
    CC._load().ccsd_default_options(C.byref(opt))
    assert (opt.mixer, opt.max_iterations, opt.energy_convergence, opt.amplitudes_convergence) == (0, 16, 1e-6, 1e-5)


def test_headers_are_plain_c():
    """The drop-in boundary is a C ABI: the three headers compile as C99 (no C++ types in the signatures)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not present")
    src = '#include "sisi4s_pt.h"\n#include "sisi4s_tn.h"\n#include "sisi4s_ccsd.h"\nint main(void){PtStats s; CcsdOptions o; (void)s; (void)o; return 0;}\n'
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                         input=src, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_host_only_entry_points():
    _ensure_built()
    lib = _lib.load()
    assert b"sm_100a" in lib.pt_version()
    assert lib.pt_num_triples(40) == 11480 and lib.pt_num_triples(5) == 35
    # partition covers the enumeration exactly, contiguously, for every rank count
    for o in (1, 2, 5, 20, 40):
        for n in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(n):
                b, e = C.c_int64(), C.c_int64()
                assert lib.pt_partition(o, n, r, C.byref(b), C.byref(e)) == 0
                assert b.value == prev and e.value >= b.value
                prev = e.value
            assert prev == lib.pt_num_triples(o)
    b, e = C.c_int64(), C.c_int64()
    assert lib.pt_partition(5, 2, 2, C.byref(b), C.byref(e)) == -1
    assert b"pt_partition" in lib.pt_last_error()
    # no C++ exception crosses the C ABI: an enumeration that cannot be allocated is an error code + message
    assert lib.pt_partition(1 << 21, 2, 0, C.byref(b), C.byref(e)) in (-1, -4)
    assert b"pt_partition" in lib.pt_last_error()


def test_partition_is_weight_balanced():
    _ensure_built()
    lib = _lib.load()
    o, n = 40, 8
    tr = [(i, j, k) for i in range(o) for j in range(i, o) for k in range(j, o)]
    w = [[6, 3, 3, 1][(i == j) + 2 * (j == k)] for (i, j, k) in tr]
    loads = []
    for r in range(n):
        b, e = C.c_int64(), C.c_int64()
        lib.pt_partition(o, n, r, C.byref(b), C.byref(e))
        loads.append(sum(w[b.value:e.value]))
    assert max(loads) <= 1.01 * (sum(w) / n)


def test_no_cpu_fallback_without_gpu():
    from conftest import gpu_available
    if gpu_available():
        pytest.skip("CUDA device present")
    _ensure_built()
    from sisi4s_b200.triples import TriplesEngine
    with pytest.raises(_lib.PtError) as ei:
        TriplesEngine(5, 19)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_every_product_entry_point_fails_loudly_without_gpu():
    """Complex and spin-orbital triples, the tensor engine and the CCSD solver: an error with a message, never
    a number computed somewhere else."""
    from conftest import gpu_available
    if gpu_available():
        pytest.skip("CUDA device present")
    _ensure_built()
    import numpy as np
    from sisi4s_b200 import synthetic as S
    from sisi4s_b200.ccsd import CcsdSolver
    from sisi4s_b200.tensor_engine import DeviceTensors, TnError
    from sisi4s_b200.triples_complex import complex_triples_energy
    from sisi4s_b200.triples_spin_orbital import spin_orbital_triples_energy
    o, v = 2, 3
    ei, ea = S.eigenenergies(o, v)
    z = lambda *shape: np.zeros(shape, order="F")
    with pytest.raises(_lib.PtError):
        spin_orbital_triples_energy(ei, ea, z(v, o), z(v, v, o, o), z(v, v, o, o), z(o, o, o, v), z(v, v, v, o))
    assert lib_message()
    with pytest.raises(_lib.PtError):
        complex_triples_energy(ei, ea, z(v, o) + 0j, z(v, v, o, o) + 0j, z(v, v, o, o) + 0j, z(v, o, o, o) + 0j,
                               np.zeros((4, o + v, o + v), dtype=complex))
    with pytest.raises(TnError):
        DeviceTensors(0)
    with pytest.raises(TnError):
        CcsdSolver(ei, ea, vertex=np.zeros((4, o + v, o + v), dtype=complex))
    # bad arguments are rejected before any device work
    with pytest.raises(ValueError):
        spin_orbital_triples_energy(ei, ea, z(v, o), z(v, v, o, o), z(v, v, o, o), z(o, o, o, v), z(v, v, o, o))


def lib_message():
    return _lib.load().pt_last_error().decode()


def test_plugin_mirror_argument_errors():
    # mirrors Algorithm::getTensorArgument / setRealArgument error behaviour
    # (reference src/algorithms/Algorithm.cxx:37-55, 364-367)
    import numpy as np
    from sisi4s_b200.triples import AlgorithmFactory, SisiException
    data = {"eps": np.zeros(3)}
    alg = AlgorithmFactory.create("CcsdPerturbativeTriples", {"HoleEigenEnergies": "$eps"}, data)
    assert alg is not None and alg.getName() == "CcsdPerturbativeTriples"
    with pytest.raises(SisiException, match="Missing argument: ParticleEigenEnergies"):
        alg.getTensorArgument("ParticleEigenEnergies")
    with pytest.raises(SisiException, match="Missing argument: CcsdPerturbativeTriplesEnergy"):
        alg.setRealArgument("CcsdPerturbativeTriplesEnergy", 1.0)
    assert AlgorithmFactory.create("NoSuchAlgorithm", {}, data) is None
    assert AlgorithmFactory.create("PerturbativeTriples", {}, data).getName() == "PerturbativeTriples"


def test_plugin_class_compiles_against_reference_headers():
    """The C++ drop-in class (CcsdPerturbativeTriplesGpu.cxx) is syntax-checked against the
    reference's OWN headers where they lie (src/algorithms/Algorithm.hpp, Data.hpp, DryTensor.hpp,
    util/*.hpp); only the absent third-party headers <ctf.hpp> and <mpi.h> are stand-ins
    (tests/stubs/).  Needs /root/reference, which exists in the build container only."""
    import shutil
    import subprocess
    ref = "/root/reference/src"
    if not os.path.isdir(ref) or shutil.which("g++") is None:
        pytest.skip("reference sources / g++ not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda_inc = "/usr/local/cuda/include"
    for src in ("CcsdPerturbativeTriplesGpu.cxx", "CcsdEnergyFromCoulombIntegralsGpu.cxx",
                "CcsdPerturbativeTriplesComplexGpu.cxx", "UPerturbativeTriplesGpu.cxx"):
        cmd = ["g++", "-std=c++17", "-fsyntax-only", "-w", "-I", os.path.join(root, "tests", "stubs"), "-I", ref,
               "-I", os.path.join(root, "include"), "-I", cuda_inc, "-I", os.path.join(root, "sisi4s_b200", "csrc"),
               os.path.join(root, "sisi4s_b200", "csrc", src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        assert res.returncode == 0, (src, res.stderr[-3000:])


def test_hole_block_walk_plan_on_the_host():
    """pt_plan_hole_blocks: the grouping pt_run uses in hole_block mode (same key / hole functions), without a
    GPU: groups tile a rank's range, at most 3b active holes, slab loads bounded by groups * 3b; the
    config-5 shape (o=100, b=6) has 969 hole-block triples."""
    _ensure_built()
    lib = _lib.load()

    def plan(o, b, begin, end):
        g, a, s = C.c_int64(), C.c_int32(), C.c_int64()
        assert lib.pt_plan_hole_blocks(o, b, begin, end, C.byref(g), C.byref(a), C.byref(s)) == 0
        return g.value, a.value, s.value

    g, a, s = plan(100, 6, 0, lib.pt_num_triples(100))
    assert g == 17 * 18 * 19 // 6 == 969 and a == 18
    assert 100 <= s <= g * 18 and s < 0.4 * g * 18          # consecutive groups share the slabs of blocks I and J
    for o, b in ((7, 2), (7, 3), (10, 4), (5, 8), (40, 8)):
        total = lib.pt_num_triples(o)
        nb = -(-o // min(b, o))
        g, a, s = plan(o, b, 0, total)
        full = nb * (nb + 1) * (nb + 2) // 6          # a one-hole last block has no triple besides the skipped i=j=k
        assert full - 1 <= g <= full and a <= min(3 * b, o) and s >= o
        # ranks' contiguous ranges: every group of the whole problem is visited by at least one rank,
        # and the per-rank walks together do not need many more launches than the whole problem
        parts = 0
        for r in range(4):
            lo, hi = C.c_int64(), C.c_int64()
            lib.pt_partition(o, 4, r, C.byref(lo), C.byref(hi))
            parts += plan(o, b, lo.value, hi.value)[0]
        assert g <= parts <= g + 3 * nb * nb
    assert lib.pt_plan_hole_blocks(5, 0, 0, 1, None, None, None) == -1


def test_partition_and_hole_block_plan_properties():
    """Property test of the two host-side planners over random shapes (hypothesis): pt_partition tiles the
    enumeration with ranks within one triple's weight of the mean; pt_plan_hole_blocks of any sub-range keeps at most
    min(3b, o) holes active, needs at least the distinct holes it touches and at most groups * 3b slab loads."""
    from hypothesis import given, settings, strategies as st
    _ensure_built()
    lib = _lib.load()

    @settings(max_examples=60, deadline=None)
    @given(o=st.integers(1, 24), n=st.integers(1, 12), b=st.integers(1, 9), cut=st.tuples(st.floats(0, 1), st.floats(0, 1)))
    def check(o, n, b, cut):
        tr = [(i, j, k) for i in range(o) for j in range(i, o) for k in range(j, o)]
        w = [[6, 3, 3, 1][(i == j) + 2 * (j == k)] for (i, j, k) in tr]
        prev, loads = 0, []
        for r in range(n):
            lo, hi = C.c_int64(), C.c_int64()
            assert lib.pt_partition(o, n, r, C.byref(lo), C.byref(hi)) == 0
            assert lo.value == prev <= hi.value
            prev = hi.value
            loads.append(sum(w[lo.value:hi.value]))
        assert prev == len(tr) and sum(loads) == sum(w)
        assert max(loads) <= sum(w) / n + 6
        lo, hi = sorted(int(c * len(tr)) for c in cut)
        g, a, s = C.c_int64(), C.c_int32(), C.c_int64()
        assert lib.pt_plan_hole_blocks(o, b, lo, hi, C.byref(g), C.byref(a), C.byref(s)) == 0
        work = [t for t in tr[lo:hi] if not (t[0] == t[1] == t[2])]
        touched = {h for t in work for h in t}
        bw = min(b, o)
        keys = {(t[0] // bw, t[1] // bw, t[2] // bw) for t in work}
        assert g.value == len(keys)
        assert a.value <= min(3 * b, o)
        assert len(touched) <= s.value <= max(1, g.value) * min(3 * b, o)
        assert (g.value == 0) == (not work)

    check()
