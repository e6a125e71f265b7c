"""Operand-feed ceiling of the fused kernel at (40,300): consumers skip LDS/DMMA (debug=1), so the
step time is what the TMA ring alone can sustain; compare with the normal run."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sisi4s_b200 import synthetic as S
from sisi4s_b200.triples import TriplesEngine
inp = S.make_inputs(40, 300, seed=2026, kind="vertex", nf=24)
with TriplesEngine(40, 300) as eng:
    eng.set_inputs(*inp.args())
    b, e = eng.partition(8, 3)
    for dbg in (0, 1):
        eng.set_option("debug", dbg)
        r = eng.run(b, e)
        print("debug", dbg, "s_kernel", r.seconds_kernel, "equiv TF/s", r.flops / r.seconds_kernel * 1e-12, flush=True)
