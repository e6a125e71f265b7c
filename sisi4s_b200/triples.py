"""Host-side mirror of the reference's plugin interface for the (T) step.

The reference is a C++ program; its plugin boundary is the ``sisi4s::Algorithm``
class (reference src/algorithms/Algorithm.hpp:26-102): an algorithm is created
from a list of named arguments, reads its inputs with ``getTensorArgument`` /
``getRealArgument``, writes its result with ``setRealArgument`` and is driven by
``run()``.  The C++ subclass that plugs the GPU library into a real sisi4s build
is sisi4s_b200/csrc/CcsdPerturbativeTriplesGpu.cxx.  Because
sisi4s itself cannot be linked in this environment (no MPI / CTF), this module
mirrors the same interface -- same algorithm names, argument keys, output keys
and error messages -- on top of the same C ABI, so tests and the benchmark read
like the reference's own YAML plans:

    - name: CcsdPerturbativeTriples            # or PerturbativeTriples
      in:  {HoleEigenEnergies: $.., ParticleEigenEnergies: $.., CcsdEnergy: $..,
            CcsdSinglesAmplitudes: $.., CcsdDoublesAmplitudes: $..,
            PPHHCoulombIntegrals: $.., HHHPCoulombIntegrals: $..,
            PPPHCoulombIntegrals: $..  (or CoulombVertex: $..)}
      out: {CcsdPerturbativeTriplesEnergy: $..}  (or PerturbativeTriplesEnergy)

All compute happens in libsisi4s_pt.so on the GPU; nothing here falls back to
the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib


# --------------------------------------------------------------------------
# thin object wrapper over the C ABI handle
# --------------------------------------------------------------------------
def _f64(a: np.ndarray) -> np.ndarray:
    """Dense FP64 column-major view/copy (CTF global layout)."""
    return np.asfortranarray(a, dtype=np.float64)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@dataclass
class RunResult:
    energy: float            # sum over the triples of this run
    per_triple: np.ndarray   # E_t for each sorted triple of the run
    seconds: float           # device time of pt_run
    seconds_kernel: float
    flops: float


class TriplesEngine:
    """One GPU's (T) engine: owns a ``pt_handle_t``."""

    def __init__(self, o: int, v: int, device: int = 0, engine: int = _lib.PT_ENGINE_FUSED,
                 keep_raw: bool = False, grid: int = 0, slab_slots: int = 0, o_all: int | None = None,
                 hole_block: int = 0, async_upload: bool = False, pin_host: bool = False, vd: int | None = None):
        """``o_all`` > ``o``: engine for a hole subset of a larger problem (pt_create_ex): ``o`` active
        holes (those of the triples that are run), ``o_all`` holes in the hole contraction.
        ``hole_block`` = b: out-of-core mode of the library (BASELINE configs[4]): the setters take the
        full tensors, T2 / PPHH / PPPH stay in host memory (kept alive by this object) and are staged
        per hole-block group of <= 3b active holes."""
        self.lib = _lib.load()
        self.o, self.v = int(o), int(v)
        self.o_all = self.o if o_all is None else int(o_all)
        self._h = C.c_void_p()
        _lib.check(self.lib.pt_create_ex(C.byref(self._h), self.o, self.o_all, self.v, int(device)))
        self._keepalive = []
        # vd: length of the particle contraction when it differs from v (stacked complex problem)
        self.vd = self.v if vd is None else int(vd)
        if self.vd != self.v:
            self.set_option("particle_contraction", self.vd)
        self.hole_block = int(hole_block)
        if hole_block:
            self.set_option("hole_block", int(hole_block))   # first: it re-dimensions the device buffers
        if async_upload:
            self.set_option("async_upload", 1)
        if pin_host:
            self.set_option("pin_host", 1)
        self.async_upload = bool(async_upload)
        self.keep_raw = bool(keep_raw)
        self.set_option("keep_raw", int(keep_raw))
        self.set_option("engine", int(engine))
        if grid:
            self.set_option("grid", int(grid))
        if slab_slots:
            self.set_option("slab_slots", int(slab_slots))

    # -- lifecycle
    def close(self):
        if self._h:
            self.lib.pt_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: int):
        _lib.check(self.lib.pt_set_option(self._h, key.encode(), int(value)))

    def sync(self):
        """Wait for the uploads / packing the setters enqueued (async_upload)."""
        _lib.check(self.lib.pt_sync(self._h))

    def _hold(self, a):
        # host buffers the library may still read after the setter returns
        if self.async_upload or self.hole_block:
            self._keepalive.append(a)
        return a

    # -- inputs
    def _shape(self, a, shape, name):
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(a.shape)}")

    def set_eigenenergies(self, epsi, epsa):
        epsi, epsa = self._hold(_f64(epsi)), self._hold(_f64(epsa))
        self._shape(epsi, (self.o,), "HoleEigenEnergies")
        self._shape(epsa, (self.v,), "ParticleEigenEnergies")
        _lib.check(self.lib.pt_set_eigenenergies(self._h, _ptr(epsi), _ptr(epsa)))

    def set_singles(self, t1):
        t1 = self._hold(_f64(t1))
        self._shape(t1, (self.v, self.o), "CcsdSinglesAmplitudes")
        _lib.check(self.lib.pt_set_singles(self._h, _ptr(t1)))

    def set_doubles(self, t2):
        t2 = self._hold(_f64(t2))
        self._shape(t2, (self.v, self.vd, self.o, self.o), "CcsdDoublesAmplitudes")
        _lib.check(self.lib.pt_set_doubles(self._h, _ptr(t2)))

    def set_singles_pair(self, t1b, vabij_b):
        """Second singles term S += 1/2 t1b (x) vabij_b (complex triples)."""
        t1b, vabij_b = self._hold(_f64(t1b)), self._hold(_f64(vabij_b))
        self._shape(t1b, (self.v, self.o), "CcsdSinglesAmplitudes (second term)")
        self._shape(vabij_b, (self.v, self.v, self.o, self.o), "PPHHCoulombIntegrals (second term)")
        _lib.check(self.lib.pt_set_singles_pair(self._h, _ptr(t1b), _ptr(vabij_b)))

    def set_doubles_hole(self, t2_xl):
        """Hole-term doubles T2[a,b,x,l], x active / l all holes (pt_create_ex engines only)."""
        t2_xl = self._hold(_f64(t2_xl))
        self._shape(t2_xl, (self.v, self.v, self.o, self.o_all), "CcsdDoublesAmplitudes (hole term)")
        _lib.check(self.lib.pt_set_doubles_hole(self._h, _ptr(t2_xl)))

    def set_pphh(self, vabij):
        vabij = self._hold(_f64(vabij))
        self._shape(vabij, (self.v, self.v, self.o, self.o), "PPHHCoulombIntegrals")
        _lib.check(self.lib.pt_set_pphh(self._h, _ptr(vabij)))

    def set_hhhp(self, vijka):
        vijka = self._hold(_f64(vijka))
        self._shape(vijka, (self.o, self.o, self.o_all, self.v), "HHHPCoulombIntegrals")
        _lib.check(self.lib.pt_set_hhhp(self._h, _ptr(vijka)))

    def set_ppph(self, vabci, slabs_per_call: int = 0):
        vabci = self._hold(_f64(vabci))
        self._shape(vabci, (self.v, self.v, self.vd, self.o), "PPPHCoulombIntegrals")
        if self.hole_block:
            return self.set_ppph_host(vabci)
        step = slabs_per_call or self.o
        slab = self.v * self.v * self.vd
        flat = vabci.reshape(-1, order="F")
        for k0 in range(0, self.o, step):
            k1 = min(self.o, k0 + step)
            part = flat[k0 * slab:k1 * slab]
            _lib.check(self.lib.pt_set_ppph_slabs(self._h, k0, k1, _ptr(part)))

    def set_ppph_host(self, vabci):
        """PPPHCoulombIntegrals kept in host memory; with ``slab_slots`` the engine uploads
        slabs on demand, so the array is kept alive by this object."""
        vabci = _f64(vabci)
        self._shape(vabci, (self.v, self.v, self.vd, self.o), "PPPHCoulombIntegrals")
        self._keepalive.append(vabci)
        _lib.check(self.lib.pt_set_ppph_host(self._h, _ptr(vabci)))

    def set_vertex(self, gamma):
        """CoulombVertex[NF,Np,Np] complex; PPPH is built on the device."""
        g = np.asarray(gamma)
        nf, np_, np2 = g.shape
        if np_ != np2:
            raise ValueError("CoulombVertex must be [NF,Np,Np]")
        gre, gim = self._hold(_f64(g.real)), self._hold(_f64(g.imag))
        _lib.check(self.lib.pt_set_vertex(self._h, nf, np_, _ptr(gre), _ptr(gim)))

    def use_vertex_integrals(self):
        """PPHH and HHHP built on the device from the resident vertex (after set_vertex) and used as
        the step's inputs: the CoulombVertex contract then needs neither tensor from the host."""
        _lib.check(self.lib.pt_use_vertex_integrals(self._h))

    def vertex_integrals(self, block: str) -> np.ndarray:
        """One real integral block ("PPHH", "HHHP", "PPPH") from the resident vertex, computed on the
        device (CoulombIntegralsFromVertex.cxx:402-403, 416-417, 430-431), column-major."""
        o, v = self.o_all, self.v
        shape = {"PPHH": (v, v, o, o), "HHHP": (o, o, o, v), "PPPH": (v, v, v, o)}[block]
        out = np.zeros(shape, dtype=np.float64, order="F")
        _lib.check(self.lib.pt_vertex_integrals(self._h, block.encode(), _ptr(out)))
        return out

    def set_inputs(self, epsi, epsa, T1, T2, Vpphh, Vhhhp, Vppph=None, vertex=None):
        """``Vpphh`` / ``Vhhhp`` None with a vertex: both are built on the device from it."""
        self.set_eigenenergies(epsi, epsa)
        self.set_singles(T1)
        self.set_doubles(T2)
        if Vpphh is not None:
            self.set_pphh(Vpphh)
        if Vhhhp is not None:
            self.set_hhhp(Vhhhp)
        if Vppph is not None:
            self.set_ppph(Vppph)
        elif vertex is not None:
            self.set_vertex(vertex)
            if Vpphh is None and Vhhhp is None:
                self.use_vertex_integrals()
        else:
            raise ValueError("Missing argument: PPPHCoulombIntegrals (or CoulombVertex)")

    # -- run
    def num_triples(self) -> int:
        return int(self.lib.pt_num_triples(self.o))

    def partition(self, nranks: int, rank: int):
        b, e = C.c_int64(), C.c_int64()
        _lib.check(self.lib.pt_partition(self.o, nranks, rank, C.byref(b), C.byref(e)))
        return int(b.value), int(e.value)

    def run(self, begin: int = 0, end: int | None = None) -> RunResult:
        end = self.num_triples() if end is None else end
        per = np.zeros(max(0, end - begin), dtype=np.float64)
        e = C.c_double(0.0)
        _lib.check(self.lib.pt_run(self._h, begin, end, C.byref(e), _ptr(per) if per.size else None))
        st = self.stats()
        return RunResult(float(e.value), per, st.seconds_run, st.seconds_kernel, st.flops_algorithmic)

    def run_list(self, triples) -> RunResult:
        """E_t of an explicit list of sorted-triple indices (enumeration over the active holes)."""
        idx = np.ascontiguousarray(triples, dtype=np.int64)
        per = np.zeros(idx.size, dtype=np.float64)
        e = C.c_double(0.0)
        _lib.check(self.lib.pt_run_list(self._h, idx.size, idx.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(e),
                                        _ptr(per) if per.size else None))
        st = self.stats()
        return RunResult(float(e.value), per, st.seconds_run, st.seconds_kernel, st.flops_algorithmic)

    def stats(self) -> _lib.PtStats:
        st = _lib.PtStats()
        _lib.check(self.lib.pt_get_stats(self._h, C.byref(st)))
        return st

    # -- debug / measurement
    def debug_w_tile(self, x, y, z, ra, rb, rc) -> np.ndarray:
        out = np.zeros(16 * 16 * 16, dtype=np.float64)
        _lib.check(self.lib.pt_debug_w_tile(self._h, x, y, z, ra, rb, rc, _ptr(out)))
        return out.reshape((16, 16, 16), order="F")

    def bench_vertex_gemm(self, what: int, reps: int = 3):
        """(seconds per build, algorithmic FLOP) of the integrals-from-vertex GEMM: 0 = one packed PPPH
        slab, 1 = the PPHH block."""
        s, f = C.c_double(), C.c_double()
        _lib.check(self.lib.pt_bench_vertex_gemm(self._h, what, reps, C.byref(s), C.byref(f)))
        return float(s.value), float(f.value)

    def bench_fp64(self, mode: int, warps_per_sm: int, iters: int):
        tf, mhz = C.c_double(), C.c_double()
        _lib.check(self.lib.pt_bench_fp64(self._h, mode, warps_per_sm, iters, C.byref(tf), C.byref(mhz)))
        return float(tf.value), float(mhz.value)


# --------------------------------------------------------------------------
# mirror of the reference plugin interface (Algorithm.hpp / Data.hpp)
# --------------------------------------------------------------------------
class SisiException(Exception):
    """Counterpart of ``throw new EXCEPTION(msg)`` (reference util/Exception.hpp:8-11)."""


class Algorithm:
    """Mirror of sisi4s::Algorithm (reference src/algorithms/Algorithm.hpp:26-102).

    ``arguments`` maps argument names to data names (the ``$Name`` symbols of a
    YAML plan, reference src/Parser.cxx:48-53) or to literal values; ``data`` is
    the global name -> value store (reference src/Data.hpp:59-74).
    """

    registry: dict[str, type] = {}
    name = "Algorithm"

    def __init__(self, arguments: dict, data: dict):
        self.arguments = dict(arguments)
        self.data = data
        self.note = ""
        self.fallible = False

    def getName(self) -> str:
        return self.name

    def isArgumentGiven(self, name: str) -> bool:
        return name in self.arguments

    def _resolve(self, name: str):
        if name not in self.arguments:
            raise SisiException(f"Missing argument: {name}")  # Algorithm.cxx:37-55
        val = self.arguments[name]
        if isinstance(val, str) and val.startswith("$"):
            key = val[1:]
            if key not in self.data:
                raise SisiException(f"Missing data: {key}")
            return self.data[key]
        return val

    def getTensorArgument(self, name: str) -> np.ndarray:
        val = self._resolve(name)
        if not isinstance(val, np.ndarray):
            raise SisiException(f"Incompatible type for argument: {name}")
        return val

    def getRealArgument(self, name: str, default=None) -> float:
        if default is not None and not self.isArgumentGiven(name):
            return float(default)
        val = self._resolve(name)
        if isinstance(val, np.ndarray):
            if val.ndim != 0 and val.size != 1:
                raise SisiException(f"Incompatible type for argument: {name}")
            return float(val.reshape(-1)[0])
        return float(val)

    def getIntegerArgument(self, name: str, default=None) -> int:
        if default is not None and not self.isArgumentGiven(name):
            return int(default)
        return int(self._resolve(name))

    def setRealArgument(self, name: str, value: float):
        if name not in self.arguments:
            raise SisiException(f"Missing argument: {name}")  # must be in the step's out: map
        target = self.arguments[name]
        key = target[1:] if isinstance(target, str) and target.startswith("$") else name
        self.data[key] = float(value)

    def run(self):
        raise NotImplementedError

    def dryRun(self):
        pass


def register(cls):
    Algorithm.registry[cls.name] = cls
    return cls


class AlgorithmFactory:
    """Mirror of AlgorithmFactory::create (reference Algorithm.hpp:104-161)."""

    @staticmethod
    def create(name: str, arguments: dict, data: dict):
        cls = Algorithm.registry.get(name)
        return cls(arguments, data) if cls else None


@register
class CcsdPerturbativeTriples(Algorithm):
    """GPU drop-in for reference src/algorithms/CcsdPerturbativeTriples.cxx:119-248.

    Accepts the compiled class's contract (``CoulombVertex`` in,
    ``CcsdPerturbativeTriplesEnergy`` out) and the PPPH contract of
    ``PerturbativeTriples`` (src/algorithms/PerturbativeTriples.cxx:172-239:
    ``PPPHCoulombIntegrals`` in, ``PerturbativeTriplesEnergy`` out).  Extra
    integer arguments: ``device`` (default 0), ``engine`` (0 fused / 1 naive).
    """

    name = "CcsdPerturbativeTriples"
    OUT_KEYS = ("CcsdPerturbativeTriplesEnergy", "PerturbativeTriplesEnergy")

    def _gather(self):
        epsi = self.getTensorArgument("HoleEigenEnergies")
        epsa = self.getTensorArgument("ParticleEigenEnergies")
        o, v = int(epsi.shape[0]), int(epsa.shape[0])
        return o, v, epsi, epsa

    def make_engine(self) -> TriplesEngine:
        o, v, epsi, epsa = self._gather()
        engine = self.getIntegerArgument("engine", _lib.PT_ENGINE_FUSED)
        # slabSlots / holeBlock: the memory options of the C++ plugin (CcsdPerturbativeTriplesGpu.cxx)
        eng = TriplesEngine(o, v, device=self.getIntegerArgument("device", 0), engine=engine,
                            keep_raw=(engine == _lib.PT_ENGINE_NAIVE),
                            slab_slots=self.getIntegerArgument("slabSlots", 0),
                            hole_block=self.getIntegerArgument("holeBlock", 0),
                            async_upload=bool(self.getIntegerArgument("asyncUpload", 1)),
                            pin_host=bool(self.getIntegerArgument("pinHost", 0)))
        eng.set_eigenenergies(epsi, epsa)
        eng.set_singles(self.getTensorArgument("CcsdSinglesAmplitudes"))
        eng.set_doubles(self.getTensorArgument("CcsdDoublesAmplitudes"))
        have_vertex = self.isArgumentGiven("CoulombVertex") and not self.isArgumentGiven("PPPHCoulombIntegrals")
        from_vertex = have_vertex and bool(self.getIntegerArgument("integralsFromVertex", 0))
        if not from_vertex:
            eng.set_pphh(self.getTensorArgument("PPHHCoulombIntegrals"))
            eng.set_hhhp(self.getTensorArgument("HHHPCoulombIntegrals"))
        if self.isArgumentGiven("PPPHCoulombIntegrals"):
            ppph = self.getTensorArgument("PPPHCoulombIntegrals")
            if eng.hole_block or self.getIntegerArgument("slabSlots", 0) or (eng.async_upload and not eng.keep_raw):
                # the tensor stays with the caller: blocked modes fetch slabs on demand; an all-resident engine
                # with asynchronous setters uploads only the slabs its triples touch, behind the first kernels
                eng.set_ppph_host(ppph)
            else:
                eng.set_ppph(ppph)
        elif have_vertex:
            eng.set_vertex(self.getTensorArgument("CoulombVertex"))
            if from_vertex:       # PPHH / HHHP on the device too (CoulombIntegralsFromVertex.cxx:402-403,416-417)
                eng.use_vertex_integrals()
        else:
            raise SisiException("Missing argument: PPPHCoulombIntegrals")
        return eng

    def run(self):
        with self.make_engine() as eng:
            res = eng.run()
            self.stats = eng.stats()
        e_triples = res.energy
        e_ccsd = self.getRealArgument("CcsdEnergy")   # mandatory, as :241 / PerturbativeTriples.cxx:229
        e = e_ccsd + e_triples
        self.log = {"e": e, "ccsd": e_ccsd, "triples": e_triples}  # LOG lines of :243-245
        given = [k for k in self.OUT_KEYS if self.isArgumentGiven(k)]
        if not given:
            raise SisiException(f"Missing argument: {self.OUT_KEYS[0]}")
        for k in given:
            self.setRealArgument(k, e)
        return e

    def dryRun(self):
        """Memory estimate.  The reference (:250-284) counts its 8 live v^3 CTF tensors plus the
        sliced inputs; the GPU step reports the device bytes per GPU instead, computed by the
        library itself (pt_estimate_device_bytes: the same formula the C++ plugin logs)."""
        o, v, _, _ = self._gather()
        slots = self.getIntegerArgument("slabSlots", 0)
        self.dry_bytes = int(_lib.load().pt_estimate_device_bytes(o, v, slots, self.getIntegerArgument("holeBlock", 0)))
        return self.dry_bytes


@register
class PerturbativeTriples(CcsdPerturbativeTriples):
    """Same step under the reference's other name (PerturbativeTriples.cxx:172-239)."""
    name = "PerturbativeTriples"


@register
class CcsdPerturbativeTriplesGpu(CcsdPerturbativeTriples):
    """Name under which the C++ plugin registers inside a sisi4s build."""
    name = "CcsdPerturbativeTriplesGpu"


def run_plan(plan: list[dict], data: dict) -> dict:
    """Minimal counterpart of Sisi4s::run (reference src/Sisi4s.cxx:22-103) for plans that
    consist of (T) steps: each entry {name, in: {...}, out: {...}}."""
    for node in plan:
        args = {}
        args.update(node.get("in", {}) or {})
        args.update(node.get("out", {}) or {})
        alg = AlgorithmFactory.create(node["name"], args, data)
        if alg is None:
            raise SisiException(f"unknown algorithm {node['name']}")
        alg.run()
    return data
