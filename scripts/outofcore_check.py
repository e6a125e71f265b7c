"""Out-of-core driver at a size that also fits resident, (40,300): one rank's share (1/world of the
hole-block groups) through sub-engines, timed end to end (host slicing + uploads + kernels), and its
per-triple energies compared with the all-resident engine.  -> gpurun_out/outofcore_check.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sisi4s_b200 import synthetic as S
from sisi4s_b200.outofcore import block_groups, deal, run_out_of_core
from sisi4s_b200.triples import TriplesEngine

o, v, block, world = (int(x) for x in (sys.argv[1:5] + ["40", "300", "8", "8"][len(sys.argv) - 1:]))
inp = S.make_inputs(o, v, seed=2026, kind="vertex", nf=24)
mine = deal(block_groups(o, block), world, 0)
idx = np.array(sorted(g for _, trip, _ in mine for g, _ in trip))
w = sum(g[2] for g in mine)
flop = 2.0 * v ** 3 * (v + o) * w
t0 = time.time()
e, per = run_out_of_core(inp.epsi, inp.epsa, inp.T1, inp.T2, inp.Vpphh, inp.Vhhhp, Vppph=inp.Vppph,
                         block=block, world=world, rank=0)
wall = time.time() - t0
with TriplesEngine(o, v) as eng:
    eng.set_inputs(*inp.args())
    t0 = time.time()
    ref = eng.run_list(idx)
    wall_res = time.time() - t0
out = {"o": o, "v": v, "block": block, "groups": len(mine), "triples": int(idx.size), "flop": flop,
       "out_of_core": {"wall_s": wall, "tflops_end_to_end": flop / wall * 1e-12, "energy": e},
       "resident": {"wall_s": wall_res, "tflops": flop / wall_res * 1e-12, "energy": ref.energy},
       "max_abs_diff_per_triple": float(np.abs(per[idx] - ref.per_triple).max())}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/outofcore_check.json", "w"), indent=1)
assert out["max_abs_diff_per_triple"] <= 1e-11
