"""Tensor file formats of the reference, for feeding the (T) step from files that a sisi4s /
cc4s run dumped.  Host-side I/O only: everything is read into
dense column-major (CTF global order) numpy arrays, which is what the C ABI takes.

* "TENS" binary   -- reference src/util/BinaryTensorFormat.hpp:9-64 (32-byte header + 8 bytes
                     per dimension) and src/util/TensorIo.cxx:119-155 (writer), :228-259 (header
                     reader), :12-36 (reader); spec docs/manual.org:181-366.  Data follow the
                     dimension headers directly, ascending global index I = a + b*N0 + ...
* text            -- src/util/TensorIo.cxx:38-117 (reader), :157-226 (writer): line 1
                     "<name> <order> <len0> ...", line 2 "<rowIndexOrder> <columnIndexOrder>",
                     then one line per row index; the stored tensor has the column indices
                     fastest, followed by the row indices.
* cc4s yaml+elements -- src/algorithms/Read.cxx:25-104,255-285: a YAML header (dimensions,
                     scalarType Real64|Complex64, elements.type TextFile|IeeeBinaryFile) next to
                     a "<stem>.elements" file with the values in ascending global index.
* eigenenergies   -- src/algorithms/cc4s/DefineHolesAndParticles.cxx:8-71: metaData.fermiEnergy
                     + metaData.energies, split into holes (<= Fermi energy) and particles.
"""
from __future__ import annotations

import os
import struct

import numpy as np

MAGIC = b"TENS"
VERSION = 0x09000
_HEADER = struct.Struct("<4si4siiiii")     # magic, version, numberType, bytesPerNumber,
                                           # numbersPerElement, order, flags, reserved
_DIM = struct.Struct("<icbh")              # length, indexName, flags, reserved


class TensorFormatError(Exception):
    """Counterpart of ``throw new EXCEPTION("Invalid file format")`` (TensorIo.cxx:236-240)."""


def write_binary(path: str, a: np.ndarray) -> None:
    """TensorIo::writeBinary (TensorIo.cxx:119-155).  Real -> (8, 1); complex -> the reference
    writes bytesPerNumber = sizeof(Complex<Real>) = 16 with numbersPerElement = 2
    (BinaryTensorFormat.hpp:47-48), kept here so the files are byte-identical."""
    a = np.asarray(a)
    cplx = np.iscomplexobj(a)
    a = a.astype(np.complex128 if cplx else np.float64, copy=False)
    with open(path, "wb") as f:
        f.write(_HEADER.pack(MAGIC, VERSION, b"IEEE", 16 if cplx else 8, 2 if cplx else 1, a.ndim, 0, 0))
        for dim, n in enumerate(a.shape):
            f.write(_DIM.pack(int(n), bytes([ord("a") + dim]), 0, 0))
        f.write(np.asfortranarray(a).tobytes(order="F"))


def read_binary_header(path: str):
    """TensorIo::readBinaryHeader (TensorIo.cxx:228-259): (shape, dtype, data offset)."""
    with open(path, "rb") as f:
        raw = f.read(_HEADER.size)
        if len(raw) < _HEADER.size:
            raise TensorFormatError("Invalid file format")
        magic, version, ntype, bpn, npe, order, flags, _ = _HEADER.unpack(raw)
        if magic != MAGIC:
            raise TensorFormatError("Invalid file format")
        if version > VERSION:
            raise TensorFormatError("Incompatible file format version")
        if ntype != b"IEEE" or flags != 0 or npe not in (1, 2):
            raise TensorFormatError("Unsupported TENS variant (need dense IEEE real/complex)")
        shape = []
        for _ in range(order):
            length, _, _, _ = _DIM.unpack(f.read(_DIM.size))
            shape.append(length)
    dtype = np.complex128 if npe == 2 else np.float64
    if (npe == 1 and bpn != 8) or (npe == 2 and bpn not in (8, 16)):
        raise TensorFormatError("Unsupported precision (need 64-bit)")
    return tuple(shape), dtype, _HEADER.size + _DIM.size * order


def read_binary(path: str, mmap: bool = False) -> np.ndarray:
    """TensorIo::readBinary (TensorIo.cxx:12-36).  ``mmap=True`` maps the file instead of
    reading it, so a PPPH tensor larger than host memory can be handed to
    ``TriplesEngine.set_ppph_host`` and paged in slab by slab."""
    if not os.path.exists(path):
        raise FileNotFoundError(f'Failed to open file "{path}"')
    shape, dtype, off = read_binary_header(path)
    count = int(np.prod(shape, dtype=np.int64)) if shape else 1
    if os.path.getsize(path) < off + count * np.dtype(dtype).itemsize:
        raise TensorFormatError("Invalid file format (truncated data)")
    if mmap:
        return np.memmap(path, dtype=dtype, mode="r", offset=off, shape=shape, order="F")
    flat = np.fromfile(path, dtype=dtype, count=count, offset=off)
    return flat.reshape(shape, order="F")


def _default_order(n):
    return "".join(chr(ord("i") + d) for d in range(n))


def write_text(path: str, a: np.ndarray, name: str = "Data", row_index_order: str = "",
               column_index_order: str = "", delimiter: str = " ") -> None:
    """TensorIo::writeText (TensorIo.cxx:157-226).  Complex tensors are written the way the reference's
    `file << values[i]` prints a std::complex: "(re,im)"."""
    a = np.asarray(a)
    cplx = np.iscomplexobj(a)
    a = a.astype(np.complex128 if cplx else np.float64)
    if cplx and "," in delimiter:
        raise TensorFormatError('complex values are written as "(re,im)": the delimiter must not contain ","')
    if row_index_order == "" and column_index_order == "":
        row_index_order = _default_order(a.ndim)
    if len(row_index_order) + len(column_index_order) != a.ndim:
        raise ValueError("Number of indices in rowIndexOrder and columnIndexOrder must match tensor order")
    stored = column_index_order + row_index_order
    b = np.transpose(a, [ord(c) - ord("i") for c in stored])       # B[stored] = A[ijk..]
    ncol = int(np.prod([a.shape[ord(c) - ord("i")] for c in column_index_order], dtype=np.int64))
    values = np.asfortranarray(b).reshape(-1, order="F")
    fmt = (lambda x: f"({x.real:.16g},{x.imag:.16g})") if cplx else (lambda x: f"{x:.16g}")
    with open(path, "w") as f:
        f.write(delimiter.join([name, str(a.ndim)] + [str(n) for n in a.shape]) + "\n")
        f.write(row_index_order + delimiter + column_index_order + "\n")
        for r in range(values.size // max(ncol, 1)):
            f.write(delimiter.join(fmt(x) for x in values[r * ncol:(r + 1) * ncol]) + "\n")


def read_text(path: str, delimiter: str = " "):
    """TensorIo::readText (TensorIo.cxx:38-117): returns (name, array in the declared order).  The
    reference splits the two header lines at blanks and scans numbers with strtod (Scanner.hpp:85-112),
    i.e. it reads blank-delimited files; here any `delimiter` a file was written with is accepted as
    well, and "(re,im)" values give a complex tensor."""
    if not os.path.exists(path):
        raise FileNotFoundError(f'Failed to open file "{path}"')
    blank = (lambda t: t.replace(delimiter, " ")) if delimiter.strip() else (lambda t: t)
    with open(path) as f:
        head = blank(f.readline()).split()
        if len(head) < 2:
            raise TensorFormatError("Invalid header line")
        name, order = head[0], int(head[1])
        lens = [int(x) for x in head[2:2 + order]]
        if len(lens) != order:
            raise TensorFormatError("Invalid header line")
        orders = blank(f.readline().rstrip("\n")).split()
        row = orders[0] if orders else ""
        col = orders[1] if len(orders) > 1 else ""
        if len(row) + len(col) != order:
            raise TensorFormatError("Number of indices in rowIndexOrder and columnIndexOrder must match tensor order")
        body = f.read()
    if "(" in body:           # NumberScanner<Complex> (Scanner.hpp:95-112)
        flat = np.array(body.replace("(", " ").replace(")", " ").replace(",", " ").split(), dtype=np.float64)
        values = flat[0::2] + 1j * flat[1::2]
    else:
        values = np.array(blank(body).split(), dtype=np.float64)
    stored = col + row
    stored_lens = [lens[ord(c) - ord("i")] for c in stored]
    if values.size != int(np.prod(stored_lens, dtype=np.int64)):
        raise TensorFormatError("Wrong number of elements read")
    b = values.reshape(stored_lens, order="F")
    a = np.transpose(b, [stored.index(c) for c in _default_order(order)])   # A[ijk..] = B[stored]
    return name, np.asfortranarray(a)


# ---- legacy FTODDUMP: Coulomb vertex + eigenenergies in one chunked binary file
#      (reference src/algorithms/CoulombVertexReader.hpp:33-51, .cxx:12-15,27-108)
FTOD_MAGIC = b"sisi4sFT"            # first 8 characters of Header::MAGIC "sisi4sFTOD" (strncmp over 8)
FTOD_REALS, FTOD_IMAGS, FTOD_EPSILONS = b"FTODreal", b"FTODimag", b"FTODepsi"
_FTOD_HEADER = struct.Struct("<8s6i")   # magic, No, Nv, NG, NSpins, kPoints, reserved_
_FTOD_CHUNK = struct.Struct("<8sq")     # magic, size (bytes of the whole chunk, its 16-byte head included)


def read_ftoddump(path: str):
    """CoulombVertexReader::run (:27-108): returns (epsi[No], epsa[Nv], Gamma[NG,Np,Np] complex).  Chunks
    may come in any order; unknown chunks are skipped, like the reference's loop does."""
    if not os.path.exists(path):
        raise FileNotFoundError("Failed to open file")
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        raw = f.read(_FTOD_HEADER.size)
        if len(raw) < _FTOD_HEADER.size:
            raise TensorFormatError("Invalid file format")
        magic, no, nv, ng, _nspins, _kpoints, _ = _FTOD_HEADER.unpack(raw)
        if magic != FTOD_MAGIC:
            raise TensorFormatError("Invalid file format")
        np_ = no + nv
        n = ng * np_ * np_
        re = im = eps = None
        offset = _FTOD_HEADER.size
        while offset < size:
            f.seek(offset)
            head = f.read(_FTOD_CHUNK.size)
            if len(head) < _FTOD_CHUNK.size:
                break
            cmagic, csize = _FTOD_CHUNK.unpack(head)
            if csize < _FTOD_CHUNK.size:
                raise TensorFormatError("Invalid chunk size")
            if cmagic == FTOD_REALS:
                re = np.fromfile(f, dtype="<f8", count=n)
            elif cmagic == FTOD_IMAGS:
                im = np.fromfile(f, dtype="<f8", count=n)
            elif cmagic == FTOD_EPSILONS:
                eps = np.fromfile(f, dtype="<f8", count=np_)
            offset += csize
    if re is None or re.size != n or eps is None or eps.size != np_:
        raise TensorFormatError("Invalid file format: vertex or eigenenergy chunk missing")
    if im is None:
        im = np.zeros(n)            # the reference leaves a missing chunk's tensor at zero
    gamma = (re + 1j * im).reshape((ng, np_, np_), order="F")
    return eps[:no].copy(), eps[no:].copy(), np.asfortranarray(gamma)


def write_ftoddump(path: str, epsi, epsa, gamma) -> None:
    """Writer of the same layout (test infrastructure / data exchange with the reference's reader)."""
    gamma = np.asarray(gamma)
    ng, np_, np2 = gamma.shape
    no, nv = len(epsi), len(epsa)
    if np_ != np2 or np_ != no + nv:
        raise ValueError("CoulombVertex must be [NG, No+Nv, No+Nv]")
    with open(path, "wb") as f:
        f.write(_FTOD_HEADER.pack(FTOD_MAGIC, no, nv, ng, 1, 1, 0))
        for magic, data in ((FTOD_REALS, np.asfortranarray(gamma.real).reshape(-1, order="F")),
                            (FTOD_IMAGS, np.asfortranarray(gamma.imag).reshape(-1, order="F")),
                            (FTOD_EPSILONS, np.concatenate([np.asarray(epsi, float), np.asarray(epsa, float)]))):
            data = np.ascontiguousarray(data, dtype="<f8")
            f.write(_FTOD_CHUNK.pack(magic, _FTOD_CHUNK.size + data.nbytes))
            f.write(data.tobytes())


def read_cc4s(yaml_path: str, mmap: bool = False) -> np.ndarray:
    """Read (reference src/algorithms/Read.cxx:25-104): YAML header + '<stem>.elements'."""
    import yaml
    with open(yaml_path) as f:
        node = yaml.safe_load(f)
    for key in ("dimensions", "elements", "scalarType", "type", "unit", "version"):
        if key not in node:
            raise TensorFormatError(f"missing key {key} in {yaml_path}")
    shape = tuple(int(d["length"]) for d in node["dimensions"])
    dtype = {"Real64": np.float64, "Complex64": np.complex128}[node["scalarType"]]
    data_path = os.path.splitext(yaml_path)[0] + ".elements"
    kind = node["elements"]["type"]
    count = int(np.prod(shape, dtype=np.int64))
    if kind == "TextFile":
        if dtype is not np.float64:
            raise TensorFormatError("text elements are real only (Read.cxx:58-76)")
        flat = np.loadtxt(data_path, dtype=np.float64, ndmin=1)
    elif kind == "IeeeBinaryFile":
        if mmap:
            return np.memmap(data_path, dtype=dtype, mode="r", shape=shape, order="F")
        flat = np.fromfile(data_path, dtype=dtype, count=count)
    else:
        raise TensorFormatError(f"unknown elements type {kind}")
    if flat.size != count:
        raise TensorFormatError("Wrong number of elements read")
    return flat.reshape(shape, order="F")


def write_cc4s(yaml_path: str, a: np.ndarray, binary: bool = True, axis_types=None) -> None:
    """Writer counterpart (header layout of Read.cxx:255-285) -- used to build test fixtures and
    to hand data to cc4s-style plans."""
    import yaml
    a = np.asarray(a)
    cplx = np.iscomplexobj(a)
    axis_types = axis_types or ["State"] * a.ndim
    node = {"version": 100, "type": "Tensor", "scalarType": "Complex64" if cplx else "Real64",
            "dimensions": [{"length": int(n), "type": t} for n, t in zip(a.shape, axis_types)],
            "elements": {"type": "IeeeBinaryFile" if binary else "TextFile"}, "unit": 1.0}
    with open(yaml_path, "w") as f:
        yaml.safe_dump(node, f)
    data_path = os.path.splitext(yaml_path)[0] + ".elements"
    flat = np.asfortranarray(a.astype(np.complex128 if cplx else np.float64)).reshape(-1, order="F")
    if binary:
        flat.tofile(data_path)
    else:
        np.savetxt(data_path, flat, fmt="%.17g")


def read_eigenenergies(yaml_path: str):
    """DefineHolesAndParticles (reference src/algorithms/cc4s/DefineHolesAndParticles.cxx:8-52):
    stable partition of metaData.energies at metaData.fermiEnergy -> (holes, particles)."""
    import yaml
    with open(yaml_path) as f:
        node = yaml.safe_load(f)
    try:
        fermi = float(node["metaData"]["fermiEnergy"])
        energies = np.array(node["metaData"]["energies"], dtype=np.float64)
    except (KeyError, TypeError) as exc:
        raise TensorFormatError(f"missing metaData.fermiEnergy / metaData.energies in {yaml_path}") from exc
    # std::partition is not stable, but for sorted spectra (the only use) the result is the same
    holes = energies[energies <= fermi]
    particles = energies[energies > fermi]
    return holes, particles
