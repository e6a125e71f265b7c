"""Tensor file formats against the byte layouts the reference defines
(BinaryTensorFormat.hpp:9-64, TensorIo.cxx, docs/manual.org:181-366) and plan parsing."""
import os
import struct

import numpy as np
import pytest

from sisi4s_b200 import tensor_io as TIO
from sisi4s_b200.plan import parse_plan, run_plan_file
from sisi4s_b200.triples import SisiException


def test_binary_header_bytes_match_the_reference_struct(tmp_path):
    a = np.arange(24, dtype=np.float64).reshape((2, 3, 4), order="F")
    p = str(tmp_path / "A.bin")
    TIO.write_binary(p, a)
    raw = open(p, "rb").read()
    # BinaryTensorHeaderBase: magic, version 0x09000, "IEEE", bytesPerNumber, numbersPerElement, order, flags, reserved
    assert raw[:4] == b"TENS" and struct.unpack("<i", raw[4:8])[0] == 0x09000 and raw[8:12] == b"IEEE"
    assert struct.unpack("<iiiii", raw[12:32]) == (8, 1, 3, 0, 0)
    # BinaryTensorDimensionHeader: length, indexName 'a'+dim, flags, reserved (8 bytes each)
    for dim, n in enumerate((2, 3, 4)):
        off = 32 + 8 * dim
        assert struct.unpack("<i", raw[off:off + 4])[0] == n and raw[off + 4:off + 5] == bytes([ord("a") + dim])
    # dense data: ascending global index I = a + b*N0 + c*N0*N1 (docs/manual.org:320)
    data = np.frombuffer(raw[32 + 24:], dtype="<f8")
    assert np.array_equal(data, np.arange(24.0))
    assert len(raw) == 32 + 24 + 24 * 8
    b = TIO.read_binary(p)
    assert b.shape == (2, 3, 4) and np.array_equal(a, b)
    m = TIO.read_binary(p, mmap=True)
    assert np.array_equal(np.asarray(m), a)


def test_binary_complex_and_errors(tmp_path):
    rng = np.random.default_rng(1)
    g = rng.normal(size=(3, 4, 4)) + 1j * rng.normal(size=(3, 4, 4))
    p = str(tmp_path / "G.bin")
    TIO.write_binary(p, g)
    raw = open(p, "rb").read()
    assert struct.unpack("<ii", raw[12:20]) == (16, 2)     # sizeof(Complex<Real>), 2 numbers per element
    assert np.array_equal(TIO.read_binary(p), g)
    bad = str(tmp_path / "bad.bin")
    open(bad, "wb").write(b"NOPE" + raw[4:])
    with pytest.raises(TIO.TensorFormatError, match="Invalid file format"):
        TIO.read_binary(bad)
    newer = str(tmp_path / "newer.bin")
    open(newer, "wb").write(raw[:4] + struct.pack("<i", 0x10000) + raw[8:])
    with pytest.raises(TIO.TensorFormatError, match="Incompatible file format version"):
        TIO.read_binary(newer)
    with pytest.raises(FileNotFoundError, match="Failed to open file"):
        TIO.read_binary(str(tmp_path / "missing.bin"))
    open(bad, "wb").write(raw[:-8])
    with pytest.raises(TIO.TensorFormatError, match="truncated"):
        TIO.read_binary(bad)


@pytest.mark.parametrize("row,col", [("", ""), ("ijk", ""), ("k", "ij"), ("ik", "j"), ("", "kji")])
def test_text_round_trip_and_layout(tmp_path, row, col):
    rng = np.random.default_rng(2)
    a = rng.normal(size=(2, 3, 4))
    p = str(tmp_path / "A.dat")
    TIO.write_text(p, a, "A", row, col)
    lines = open(p).read().splitlines()
    assert lines[0] == "A 3 2 3 4"
    name, b = TIO.read_text(p)
    assert name == "A" and b.shape == a.shape and np.allclose(a, b, rtol=0, atol=1e-15 * np.abs(a).max() * 10)
    if (row, col) == ("k", "ij"):
        # one line per k, columns run over (i fastest, j)
        assert len(lines) == 2 + 4
        first = np.array(lines[2].split(), dtype=float)
        assert np.allclose(first, a[:, :, 0].reshape(-1, order="F"), rtol=1e-15)


def test_cc4s_yaml_elements_and_eigenenergies(tmp_path):
    rng = np.random.default_rng(3)
    g = rng.normal(size=(5, 6, 6)) + 1j * rng.normal(size=(5, 6, 6))
    yp = str(tmp_path / "CoulombVertex.yaml")
    TIO.write_cc4s(yp, g, binary=True, axis_types=["AuxiliaryField", "State", "State"])
    assert os.path.exists(str(tmp_path / "CoulombVertex.elements"))
    assert np.array_equal(TIO.read_cc4s(yp), g)
    r = rng.normal(size=(4, 3))
    rp = str(tmp_path / "R.yaml")
    TIO.write_cc4s(rp, r, binary=False)
    assert np.allclose(TIO.read_cc4s(rp), r, rtol=1e-15)
    ep = str(tmp_path / "EigenEnergies.yaml")
    open(ep, "w").write("metaData:\n  fermiEnergy: 0.0\n  energies: [-1.5, -0.7, -0.2, 0.3, 0.9, 1.4, 2.2]\n")
    holes, parts = TIO.read_eigenenergies(ep)
    assert np.array_equal(holes, [-1.5, -0.7, -0.2]) and np.array_equal(parts, [0.3, 0.9, 1.4, 2.2])


def test_plan_parsing_and_file_algorithms(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    a = np.random.default_rng(4).normal(size=(3, 2))
    TIO.write_binary("A.bin", a)
    open("in.yaml", "w").write("""
- name: Nop
  tol: &tol 1e-8
- name: TensorReader
  in: {file: "A.bin", mode: "binary"}
  out: {Data: $B}
- name: TensorWriter
  in: {Data: $B, rowIndexOrder: "j", columnIndexOrder: "i"}
- {name: TensorReader, in: {file: "B.dat"}, out: {Data: $C}}
""")
    data = run_plan_file("in.yaml", log=lambda *_: None)
    assert np.array_equal(data["B"], a) and np.allclose(data["C"], a, rtol=1e-15)
    with pytest.raises(SisiException, match="not provided"):
        open("bad.yaml", "w").write("- name: HartreeFockFromGaussian\n  in: {}\n")
        run_plan_file("bad.yaml", log=lambda *_: None)
    assert [n["name"] for n in parse_plan(open("in.yaml").read())] == ["TensorReader", "TensorWriter", "TensorReader"]


def test_ueg_vertex_generator_step():
    """UegVertexGenerator as a plan step: closed-shell checks and outputs of the reference algorithm."""
    from sisi4s_b200.triples import AlgorithmFactory
    data = {}
    args = dict(No=7, Nv=26, rs=1.0, CoulombVertex="$CoulombVertex", HoleEigenEnergies="$HoleEigenEnergies",
                ParticleEigenEnergies="$ParticleEigenEnergies")
    AlgorithmFactory.create("UegVertexGenerator", args, data).run()
    assert data["CoulombVertex"].shape == (257, 33, 33) and data["HoleEigenEnergies"].shape == (7,)
    assert data["HoleEigenEnergies"].max() < data["ParticleEigenEnergies"].min()
    with pytest.raises(ValueError, match="closed shells"):
        AlgorithmFactory.create("UegVertexGenerator", dict(args, No=5), data).run()
    with pytest.raises(SisiException, match="Invalid rs"):
        AlgorithmFactory.create("UegVertexGenerator", dict(args, rs=0.0), data).run()


def test_text_format_complex_and_delimiters(tmp_path):
    """Advisor r01: complex tensors keep their imaginary part ("(re,im)", the reference's operator<< of
    std::complex, read back by NumberScanner<Complex>), and a file written with another delimiter
    reads back."""
    rng = np.random.default_rng(5)
    g = np.asfortranarray(rng.standard_normal((3, 4, 2)) + 1j * rng.standard_normal((3, 4, 2)))
    p = str(tmp_path / "CoulombVertex.dat")
    TIO.write_text(p, g, "CoulombVertex")
    name, back = TIO.read_text(p)
    assert name == "CoulombVertex" and np.iscomplexobj(back) and np.abs(back - g).max() < 1e-15
    a = np.asfortranarray(rng.standard_normal((4, 5)))
    TIO.write_text(p, a, "A", row_index_order="i", column_index_order="j", delimiter=",")
    name, back = TIO.read_text(p, delimiter=",")
    assert name == "A" and np.abs(back - a).max() < 1e-15
    with pytest.raises(TIO.TensorFormatError):
        TIO.write_text(p, g, "G", delimiter=",")


def test_ftoddump_reader_follows_the_reference_layout(tmp_path):
    """Legacy FTODDUMP (CoulombVertexReader.hpp:33-51): 32-byte header `sisi4sFT` + 6 int32, chunks
    `FTODreal` / `FTODimag` / `FTODepsi` with their total size in bytes, dense column-major doubles."""
    import struct
    from sisi4s_b200 import synthetic as S
    from sisi4s_b200.plan import run_plan_file
    no, nv, ng = 3, 5, 7
    gamma = S.make_vertex(no, nv, seed=17, nf=ng)
    epsi, epsa = S.eigenenergies(no, nv)
    p = str(tmp_path / "FTODDUMP")
    TIO.write_ftoddump(p, epsi, epsa, gamma)
    raw = open(p, "rb").read()
    assert raw[:8] == b"sisi4sFT" and struct.unpack("<6i", raw[8:32]) == (no, nv, ng, 1, 1, 0)
    n = ng * (no + nv) ** 2
    assert raw[32:40] == b"FTODreal" and struct.unpack("<q", raw[40:48])[0] == 16 + 8 * n
    assert np.frombuffer(raw[48:48 + 8 * n]).tolist() == gamma.real.reshape(-1, order="F").tolist()
    assert len(raw) == 32 + 2 * (16 + 8 * n) + 16 + 8 * (no + nv)
    e1, e2, g = TIO.read_ftoddump(p)
    assert np.array_equal(e1, epsi) and np.array_equal(e2, epsa) and np.array_equal(g, gamma)
    # chunk order does not matter, unknown chunks are skipped (the reference's while loop, :83-100)
    chunks = [raw[32:48 + 8 * n], raw[48 + 8 * n:64 + 16 * n], raw[64 + 16 * n:]]
    odd = b"FTODxxxx" + struct.pack("<q", 16 + 24) + b"\0" * 24
    open(p, "wb").write(raw[:32] + chunks[2] + odd + chunks[1] + chunks[0])
    e1, e2, g = TIO.read_ftoddump(p)
    assert np.array_equal(e1, epsi) and np.array_equal(g, gamma)
    # as a plan step
    open(tmp_path / "in.yaml", "w").write(f"""
- name: CoulombVertexReader
  in: {{file: "{p}"}}
  out: {{CoulombVertex: $CoulombVertex, HoleEigenEnergies: $HoleEigenEnergies, ParticleEigenEnergies: $ParticleEigenEnergies}}
""")
    data = run_plan_file(str(tmp_path / "in.yaml"), log=lambda *_: None)
    assert np.array_equal(data["CoulombVertex"], gamma) and np.array_equal(data["ParticleEigenEnergies"], epsa)
    open(p, "wb").write(b"notafile" + raw[8:])
    with pytest.raises(TIO.TensorFormatError, match="Invalid file format"):
        TIO.read_ftoddump(p)


def test_file_headers_match_the_structs_of_the_reference_headers(tmp_path):
    """oracle/_ref/file_formats.txt is printed by a program compiled against the reference's OWN
    src/util/BinaryTensorFormat.hpp and src/algorithms/CoulombVertexReader.hpp (oracle/ref_format_dump.cxx):
    the bytes of a TENS header / dimension headers and the layout of the FTODDUMP structs.  The writers here
    must produce exactly those bytes."""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "file_formats.txt")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    ref = {}
    for line in open(path):
        tag, *rest = line.split()
        ref[tag] = rest
    hexbytes = lambda tag: bytes(int(x, 16) for x in ref[tag])
    p = str(tmp_path / "t.bin")
    TIO.write_binary(p, np.zeros((5, 7, 3, 2), order="F"))
    raw = open(p, "rb").read()
    assert raw[:32] == hexbytes("tens_real_order4")
    for dim in range(4):
        assert raw[32 + 8 * dim:40 + 8 * dim] == hexbytes(f"tens_dim{dim}")
    assert len(raw) == 32 + 4 * 8 + 8 * 5 * 7 * 3 * 2
    TIO.write_binary(p, np.zeros((2, 3, 4), dtype=complex, order="F"))
    assert open(p, "rb").read()[:32] == hexbytes("tens_complex_order3")
    # the first line reads "sizeof_header 32 sizeof_dim 8"
    assert ref["sizeof_header"][0] == "32" and ref["sizeof_header"][2] == "8"
    # FTODDUMP structs: Header = magic[8] + 6 int32, Chunk = magic[8] + int64 size
    f = dict(zip(ref["ftod_header"][0::2], (int(x) for x in ref["ftod_header"][1::2])))
    assert f == {"size": TIO._FTOD_HEADER.size, "magic": 0, "No": 8, "Nv": 12, "NG": 16, "NSpins": 20, "kPoints": 24, "reserved": 28}
    c = dict(zip(ref["ftod_chunk"][0::2], (int(x) for x in ref["ftod_chunk"][1::2])))
    assert c == {"size": TIO._FTOD_CHUNK.size, "magic": 0, "size_field": 8, "magic_len": 8}
