"""Quick timing of the fused kernel at (20,100) full and (40,300) partial ranges."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sisi4s_b200 import synthetic as S
from sisi4s_b200.triples import TriplesEngine

out = {}
inp = S.make_inputs(20, 100, seed=2026, kind="vertex")
with TriplesEngine(20, 100) as eng:
    eng.set_inputs(*inp.args())
    for rep in range(3):
        r = eng.run()
    out["o20_v100"] = {"E": r.energy, "s": r.seconds, "s_kernel": r.seconds_kernel, "tflops": r.flops / r.seconds_kernel * 1e-12}
print(json.dumps(out), flush=True)
t0 = time.time()
inp = S.make_inputs(40, 300, seed=2026, kind="vertex", nf=24)
print("gen s", time.time() - t0, flush=True)
with TriplesEngine(40, 300) as eng:
    t0 = time.time(); eng.set_inputs(*inp.args()); st = eng.stats()
    print("upload+pack wall", time.time() - t0, "dev", st.seconds_upload, "GB", st.bytes_h2d / 1e9, flush=True)
    for (b, e) in ((1000, 1040), (1000, 1040), (5000, 5100), (5000, 5200)):
        r = eng.run(b, e)
        out[f"o40_v300_{b}_{e}"] = {"E": r.energy, "s_kernel": r.seconds_kernel, "tflops": r.flops / r.seconds_kernel * 1e-12}
        print(json.dumps(out[f"o40_v300_{b}_{e}"]), flush=True)
json.dump(out, open("gpurun_out/quick_timing.json", "w"), indent=1)
