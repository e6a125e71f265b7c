"""Quick timing of the fused kernel: (20,100) full and one bench step (1/8 of the sorted
triples) at (40,300), for each "key=value,key=value" option set given on the command line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sisi4s_b200 import synthetic as S
from sisi4s_b200.triples import TriplesEngine

variants = sys.argv[1:] or ["order=1"]
def apply(eng, var):
    for kv in var.split(","):
        k, v = kv.split("=")
        eng.set_option(k, int(v))
out = {}
inp = S.make_inputs(20, 100, seed=2026, kind="vertex")
with TriplesEngine(20, 100) as eng:
    eng.set_inputs(*inp.args())
    for var in variants:
        apply(eng, var)
        for rep in range(3):
            r = eng.run()
        out[f"o20_v100 {var}"] = {"E": r.energy, "s_kernel": r.seconds_kernel, "tflops": r.flops / r.seconds_kernel * 1e-12}
        print(var, json.dumps(out[f"o20_v100 {var}"]), flush=True)
t0 = time.time()
inp = S.make_inputs(40, 300, seed=2026, kind="vertex", nf=24)
print("gen s", time.time() - t0, flush=True)
with TriplesEngine(40, 300) as eng:
    t0 = time.time(); eng.set_inputs(*inp.args()); st = eng.stats()
    print("upload+pack wall", time.time() - t0, "dev", st.seconds_upload, "GB", st.bytes_h2d / 1e9, flush=True)
    b, e = eng.partition(8, 3)
    eng.run(b, e)
    for var in variants:
        apply(eng, var)
        for part in (3, 6):
            b, e = eng.partition(8, part)
            r = eng.run(b, e)
            key = f"o40_v300 {var} part{part}"
            out[key] = {"E": r.energy, "s_kernel": r.seconds_kernel, "tflops": r.flops / r.seconds_kernel * 1e-12}
            print(key, json.dumps(out[key]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/quick_timing.json", "w"), indent=1)
