"""Device CCSD solver (sisi4s_b200/ccsd.py on the tensor engine) at a moderate synthetic size: seconds per
iteration, GEMM TFLOP/s of the contractions, energy vs the literal NumPy oracle for the first iterations.
    python scripts/ccsd_timing.py [o v nf iters] -> gpurun_out/ccsd_timing.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sisi4s_b200 import synthetic as S
from sisi4s_b200.ccsd import CcsdSolver
o, v, nf, iters = (int(x) for x in (sys.argv[1:5] + ["16", "96", "64", "6"][len(sys.argv) - 1:]))
epsi, epsa = S.eigenenergies(o, v)
gamma = S.make_vertex(o, v, seed=4, nf=nf, kappa=0.35 * S.default_kappa(o, v, nf) / S.default_kappa(3, 6, 14))
np_ = o + v
h, p = slice(0, o), slice(np_ - v, np_)
out = {"o": o, "v": v, "nf": nf}
t0 = time.time()
with CcsdSolver(epsi, epsa, vertex=gamma) as s:       # the six integral blocks are built on the device
    V = {b: s.get_integrals(b) for b in ("PPHH", "PHPH", "HHHH", "HHHP", "PPPH", "PPPP")} if v <= 40 else None
    out["integrals_s"] = time.time() - t0
    s.solve(mixer="DiisMixer", max_iterations=2)          # warm-up: workspace allocation, first launches
    s.set_amplitudes(np.zeros((v, o)), np.zeros((v, v, o, o)))
    t0 = time.time()
    res = s.solve(mixer="DiisMixer", max_iterations=iters, energy_convergence=1e-14, amplitudes_convergence=1e-14)
    dt = time.time() - t0
    out.update(iterations=res["iterations"], seconds=dt, s_per_iteration=dt / res["iterations"], energy=res["energy"],
               gemm_tflops=res["stats"]["flops"] / dt * 1e-12, kernel_launches=res["stats"]["launches"])
if v <= 40:
    from oracle import ccsd_ref as R
    ref = R.solve(epsi, epsa, V, mixer="DiisMixer", max_iterations=iters, energy_convergence=1e-14, amplitudes_convergence=1e-14)
    out["energy_oracle"] = ref["energy"]
    out["abs_diff"] = abs(ref["energy"] - res["energy"])
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/ccsd_timing_o{o}_v{v}.json", "w"), indent=1)
