"""YAML execution plans for the (T) step: the reference's file-format algorithms next to the
GPU step, so that a plan written for sisi4s (``sisi4s -i in.yaml``; parser: reference
src/Parser.cxx:22-109, driver: src/Sisi4s.cxx:22-103) runs unchanged as far as it consists of
readers/writers and the triples step:

    - name: DefineHolesAndParticles   {in: {fileName: EigenEnergies.yaml}, out: {...}}
    - name: Read                      {in: {fileName: CoulombVertex.yaml}, out: {destination: $CoulombVertex}}
    - name: TensorReader              {in: {file: T2.bin, mode: binary}, out: {Data: $CcsdDoublesAmplitudes}}
    - name: CcsdPerturbativeTriples   {in: {...}, out: {CcsdPerturbativeTriplesEnergy: $E}}
    - name: TensorWriter              {in: {Data: $E}}
    (also: UegVertexGenerator {in: {No, Nv, rs}, out: {CoulombVertex, HoleEigenEnergies, ParticleEigenEnergies}})

    python -m sisi4s_b200 in.yaml        (cwd-relative file names, like the reference)

The steps next to the path are registered further down: UegVertexGenerator,
CoulombIntegralsFromVertex (N1) and CcsdEnergyFromCoulombIntegrals[Reference] (N3).  Hartree-Fock and
the Gaussian integral engines stay out of scope; a plan naming an unknown algorithm fails with the
reference's behaviour for an unknown name made explicit.
"""
from __future__ import annotations

import numpy as np

from . import tensor_io as TIO
from .triples import Algorithm, AlgorithmFactory, SisiException, register


def _data_name(alg: Algorithm, arg: str) -> str:
    """getArgumentData(arg)->getName() (Algorithm.cxx:57-76): the symbol behind ``$Name``."""
    if arg not in alg.arguments:
        raise SisiException(f"Missing argument: {arg}")
    val = alg.arguments[arg]
    return val[1:] if isinstance(val, str) and val.startswith("$") else arg


def _text(alg: Algorithm, name: str, default=None) -> str:
    """getTextArgument (Algorithm.cxx:78-97)."""
    if name not in alg.arguments:
        if default is None:
            raise SisiException(f"Missing argument: {name}")
        return default
    return str(alg.arguments[name])


@register
class TensorReader(Algorithm):
    """reference src/algorithms/TensorReader.cxx:14-68 (modes "text" | "binary")."""
    name = "TensorReader"

    def run(self):
        name = _data_name(self, "Data")
        mode = _text(self, "mode", "text")
        if self.getIntegerArgument("precision", 64) != 64:
            raise SisiException("TensorReader: only precision 64 is supported")
        if mode == "binary":
            a = TIO.read_binary(_text(self, "file", name + ".bin"),
                                mmap=bool(self.getIntegerArgument("mmap", 0)))
        else:
            _, a = TIO.read_text(_text(self, "file", name + ".dat"), _text(self, "delimiter", " "))
        self.data[name] = a


@register
class ComplexTensorReader(TensorReader):
    """reference src/algorithms/ComplexTensorReader.cxx: same formats, complex elements."""
    name = "ComplexTensorReader"


@register
class TensorWriter(Algorithm):
    """reference src/algorithms/TensorWriter.cxx:20-60."""
    name = "TensorWriter"

    def run(self):
        name = _data_name(self, "Data")
        val = self._resolve("Data")
        a = np.asarray(val, dtype=np.complex128 if np.iscomplexobj(val) else np.float64)
        binary = _text(self, "mode", "text") == "binary"
        path = _text(self, "file", name + (".bin" if binary else ".dat"))
        if binary:
            TIO.write_binary(path, a)
        else:
            TIO.write_text(path, a, name, _text(self, "rowIndexOrder", ""),
                           _text(self, "columnIndexOrder", ""), _text(self, "delimiter", " "))


@register
class Read(Algorithm):
    """reference src/algorithms/Read.cxx:25-104 (cc4s yaml + .elements)."""
    name = "Read"

    def run(self):
        self.data[_data_name(self, "destination")] = TIO.read_cc4s(_text(self, "fileName"))


@register
class DefineHolesAndParticles(Algorithm):
    """reference src/algorithms/cc4s/DefineHolesAndParticles.cxx:8-52."""
    name = "DefineHolesAndParticles"

    def run(self):
        holes, particles = TIO.read_eigenenergies(_text(self, "fileName"))
        self.data[_data_name(self, "HoleEigenEnergies")] = holes
        self.data[_data_name(self, "ParticleEigenEnergies")] = particles


@register
class CoulombVertexReader(Algorithm):
    """Legacy FTODDUMP reader (reference src/algorithms/CoulombVertexReader.cxx:27-108): `file` ->
    HoleEigenEnergies, ParticleEigenEnergies, CoulombVertex.  `unrestricted: 1` (:172-180) is not
    supported (the (T) step here is closed shell)."""
    name = "CoulombVertexReader"

    def run(self):
        if self.getIntegerArgument("unrestricted", 0):
            raise SisiException("CoulombVertexReader: unrestricted vertices are not supported")
        try:
            epsi, epsa, gamma = TIO.read_ftoddump(_text(self, "file"))
        except FileNotFoundError:
            raise SisiException("Failed to open file")
        except TIO.TensorFormatError as e:
            raise SisiException(str(e))
        for key, val in (("HoleEigenEnergies", epsi), ("ParticleEigenEnergies", epsa), ("CoulombVertex", gamma)):
            if self.isArgumentGiven(key):
                self.data[_data_name(self, key)] = val


@register
class UegVertexGenerator(Algorithm):
    """Counterpart of reference src/algorithms/UegVertexGenerator.cxx:51-229 (arguments No, Nv, rs;
    outputs CoulombVertex, HoleEigenEnergies, ParticleEigenEnergies).  The vertex is written in real
    standing-wave orbitals (sisi4s_b200/ueg.py), for which the reference's real-integral formulas
    are exact; the reference's own generator writes plane waves "just for profiling" (:148-149).
    NF and halfGrid are accepted and ignored (the full momentum-transfer grid is always used)."""
    name = "UegVertexGenerator"

    def run(self):
        from . import ueg
        no, nv = self.getIntegerArgument("No"), self.getIntegerArgument("Nv")
        rs = self.getRealArgument("rs")
        if no <= 0:
            raise SisiException("No larger zero please")          # :63
        if rs <= 0.0:
            raise SisiException("Invalid rs")                     # :64
        epsi, epsa, gamma = ueg.make_ueg(no, nv, rs)
        self.data[_data_name(self, "CoulombVertex")] = gamma
        self.data[_data_name(self, "HoleEigenEnergies")] = epsi
        self.data[_data_name(self, "ParticleEigenEnergies")] = epsa


@register
class CoulombIntegralsFromVertex(Algorithm):
    """Real Coulomb integral blocks from the vertex ON THE DEVICE, reference src/algorithms/
    CoulombIntegralsFromVertex.cxx:390-433 (directly computed blocks, V = Re.Re + Im.Im with the
    reference's index strings) and :436-560 (blocks derived by index permutation).  Every statement is
    one contraction / permuted copy of the device tensor engine (tensor_engine.DeviceTensors: the
    library's FP64 tensor-core GEMM).  `complex: 1` and `antisymmetrize: 1` are not supported.  Only the
    blocks named in the step's `out:` map are built."""
    name = "CoulombIntegralsFromVertex"
    # name -> (first vertex part, its indices, second part, its indices, output indices)
    DIRECT = {"PPHHCoulombIntegrals": ("ai", "ai", "ai", "bj", "abij"),      # :402-403
              "HHHHCoulombIntegrals": ("ij", "ik", "ij", "jl", "ijkl"),      # :409-410
              "HHHPCoulombIntegrals": ("ij", "ik", "ai", "aj", "ijka"),      # :416-417
              "PPPPCoulombIntegrals": ("ab", "ac", "ab", "bd", "abcd"),      # :423-424
              "PPPHCoulombIntegrals": ("ab", "ac", "ai", "bi", "abci"),      # :430-431
              "PHPHCoulombIntegrals": ("ab", "ab", "ij", "ij", "aibj")}      # :395-396
    # name -> (source block, source index string, output index string)
    DERIVED = {"HPPHCoulombIntegrals": ("PPHHCoulombIntegrals", "baij", "iabj"),   # :443
               "HPHPCoulombIntegrals": ("PHPHCoulombIntegrals", "aibj", "iajb"),   # :455
               "HPPPCoulombIntegrals": ("PPPHCoulombIntegrals", "cbai", "iabc"),   # :467
               "HHPHCoulombIntegrals": ("HHHPCoulombIntegrals", "jika", "ijak"),   # :479
               "HHPPCoulombIntegrals": ("PPHHCoulombIntegrals", "abij", "ijab"),   # :491
               "PPHPCoulombIntegrals": ("PPPHCoulombIntegrals", "baci", "abic"),   # :503
               "PHHPCoulombIntegrals": ("PPHHCoulombIntegrals", "abji", "aijb"),   # :515
               "PHPPCoulombIntegrals": ("PPPHCoulombIntegrals", "bcai", "aibc"),   # :527
               "PHHHCoulombIntegrals": ("HHHPCoulombIntegrals", "kjia", "aijk"),   # :539
               "HPHHCoulombIntegrals": ("HHHPCoulombIntegrals", "jkia", "iajk")}   # :551

    def run(self):
        from .tensor_engine import DeviceTensors
        if self.getIntegerArgument("complex", 0) or self.getIntegerArgument("antisymmetrize", 0):
            raise SisiException("CoulombIntegralsFromVertex: only real, non-antisymmetrized integrals")
        g = self.getTensorArgument("CoulombVertex")
        no = int(self.getTensorArgument("HoleEigenEnergies").shape[0])
        nv = int(self.getTensorArgument("ParticleEigenEnergies").shape[0])
        np_ = g.shape[1]
        h, pt = slice(0, no), slice(np_ - nv, np_)                       # particles = last Nv states (:121-136)
        extent = lambda letters: tuple(nv if c in "abcd" else no for c in letters)
        wanted = [n for n in list(self.DIRECT) + list(self.DERIVED) if self.isArgumentGiven(n)]
        with DeviceTensors(self.getIntegerArgument("device", 0)) as eng:
            parts = {}
            for key, blk in (("ij", g[:, h, h]), ("ai", g[:, pt, h]), ("ab", g[:, pt, pt])):
                parts[key] = (eng.tensor(blk.shape, blk.real), eng.tensor(blk.shape, blk.imag))   # fromComplexTensor
            cache = {}

            def direct(name):
                if name not in cache:
                    p1, i1, p2, i2, out = self.DIRECT[name]
                    t = eng.tensor(extent(out))
                    for part, beta in ((0, 0.0), (1, 1.0)):              # Re.Re, then += Im.Im
                        eng.contract(1.0, parts[p1][part], "G" + i1, parts[p2][part], "G" + i2, beta, t, out)
                    cache[name] = t
                return cache[name]

            for name in wanted:
                if name in self.DIRECT:
                    val = direct(name).get()
                else:
                    src, si, so = self.DERIVED[name]
                    t = eng.tensor(extent(so))
                    eng.add(1.0, direct(src), si, 0.0, t, so)
                    val = t.get()
                    t.free()
                self.data[_data_name(self, name)] = val


@register
class CcsdEnergyFromCoulombIntegralsReference(Algorithm):
    """CCSD step in front of the triples: reference CcsdEnergyFromCoulombIntegralsReference.cxx:29-295
    driven by ClusterSinglesDoublesAlgorithm.cxx:37-128, on the device (sisi4s_b200/ccsd.py).  Same
    arguments as the reference: the integral blocks PPHH, PHPH, HHHH, HHHP, PPPH, PPPP, eigenenergies,
    `mixer` (LinearMixer default / DiisMixer), `maxResidua`, `mixingRatio`, `maxIterations`,
    `energyConvergence`, `amplitudesConvergence` (relative criteria), `levelShift`; outputs CcsdEnergy,
    CcsdSinglesAmplitudes, CcsdDoublesAmplitudes.  Blocks that are missing are built from `CoulombVertex`
    when the plan passes it.  Not converging is a logged WARNING, as in the reference (:120-124)."""
    name = "CcsdEnergyFromCoulombIntegralsReference"

    def run(self):
        from . import ccsd
        epsi = self.getTensorArgument("HoleEigenEnergies")
        epsa = self.getTensorArgument("ParticleEigenEnergies")
        V, missing = {}, []
        for b in ccsd.BLOCKS:
            key = b + "CoulombIntegrals"
            if self.isArgumentGiven(key):
                V[b] = self.getTensorArgument(key)
            else:
                missing.append(key)
        vertex = None
        if missing:
            # blocks the plan does not pass are built on the device from the vertex, when that is given
            if not self.isArgumentGiven("CoulombVertex"):
                raise SisiException(f"Missing argument: {missing[0]}")
            vertex = self.getTensorArgument("CoulombVertex")
        mixer = _text(self, "mixer", "LinearMixer")
        if mixer not in ccsd.MIXERS:
            raise SisiException(f"Mixer not implemented: {mixer}")        # ClusterSinglesDoublesAlgorithm.cxx:50-54
        with ccsd.CcsdSolver(epsi, epsa, V, device=self.getIntegerArgument("device", 0), vertex=vertex) as solver:
            if self.isArgumentGiven("initialSinglesAmplitudes") or self.isArgumentGiven("initialDoublesAmplitudes"):
                solver.set_amplitudes(
                    self.getTensorArgument("initialSinglesAmplitudes") if self.isArgumentGiven("initialSinglesAmplitudes") else None,
                    self.getTensorArgument("initialDoublesAmplitudes") if self.isArgumentGiven("initialDoublesAmplitudes") else None)
            res = solver.solve(mixer=mixer, max_residua=self.getIntegerArgument("maxResidua", 4),
                               mixing_ratio=self.getRealArgument("mixingRatio", 1.0),
                               max_iterations=self.getIntegerArgument("maxIterations", ccsd.DEFAULT_MAX_ITERATIONS),
                               energy_convergence=self.getRealArgument("energyConvergence", ccsd.DEFAULT_ENERGY_CONVERGENCE),
                               amplitudes_convergence=self.getRealArgument("amplitudesConvergence", ccsd.DEFAULT_AMPLITUDES_CONVERGENCE),
                               level_shift=self.getRealArgument("levelShift", ccsd.DEFAULT_LEVEL_SHIFT))
        self.log = {"e": res["energy"], "dir": res["direct"], "exc": res["exchange"]}
        self.iterations = res["iterations"]
        self.converged = res["converged"]
        if not res["converged"]:
            self.note = "WARNING: energy or amplitudes convergence not reached."
        self.setRealArgument("CcsdEnergy", res["energy"])
        for key, val in (("CcsdSinglesAmplitudes", res["T1"]), ("CcsdDoublesAmplitudes", res["T2"])):
            if self.isArgumentGiven(key):
                self.data[_data_name(self, key)] = val


@register
class CcsdEnergyFromCoulombIntegrals(CcsdEnergyFromCoulombIntegralsReference):
    """Same step under the reference's other name."""
    name = "CcsdEnergyFromCoulombIntegrals"


def _load_yaml12(text: str):
    """yaml-cpp (the reference's parser) follows YAML 1.2: only true/false are booleans.  PyYAML's
    YAML 1.1 resolver would turn the argument key `No` of UegVertexGenerator into False."""
    import re
    import yaml

    class Loader(yaml.SafeLoader):
        pass

    Loader.yaml_implicit_resolvers = {k: [(tag, rx) for tag, rx in v if tag != "tag:yaml.org,2002:bool"]
                                      for k, v in yaml.SafeLoader.yaml_implicit_resolvers.items()}
    Loader.add_implicit_resolver("tag:yaml.org,2002:bool", re.compile(r"^(?:true|True|TRUE|false|False|FALSE)$"),
                                 list("tTfF"))
    return yaml.load(text, Loader=Loader)


def parse_plan(text: str) -> list[dict]:
    """Parser::parse (reference src/Parser.cxx:22-109): a YAML sequence of
    {name, in: {..}, out: {..}} nodes; other keys of a node (anchors on a Nop step) are ignored."""
    nodes = _load_yaml12(text)
    if not isinstance(nodes, list):
        raise SisiException("the execution plan must be a YAML sequence of algorithms")
    plan = []
    for n in nodes:
        if not isinstance(n, dict) or "name" not in n:
            raise SisiException("every step needs a name")
        if n["name"] == "Nop":
            continue
        plan.append({"name": n["name"], "in": n.get("in") or {}, "out": n.get("out") or {}})
    return plan


def run_plan_file(path: str, data: dict | None = None, log=print) -> dict:
    """Sisi4s::run (reference src/Sisi4s.cxx:22-103) for the supported algorithms."""
    data = {} if data is None else data
    with open(path) as f:
        plan = parse_plan(f.read())
    for n, node in enumerate(plan):
        args = dict(node["in"])
        args.update(node["out"])
        alg = AlgorithmFactory.create(node["name"], args, data)
        if alg is None:
            raise SisiException(f"step {n + 1}: algorithm {node['name']} is not provided by sisi4s_b200 "
                                f"(available: {sorted(Algorithm.registry)})")
        log(f"step={n + 1} {node['name']}")
        alg.run()
        if getattr(alg, "log", None):
            for k, v in alg.log.items():
                log(f"  {k}={v:.15g}")
    return data


from . import triples_complex, triples_spin_orbital  # noqa: E402,F401  (register CcsdPerturbativeTriplesComplex, UPerturbativeTriples)
