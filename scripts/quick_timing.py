"""Quick timing of the fused kernel: (20,100) full and one bench step (1/8 of the sorted
triples) at (40,300), for each item order given on the command line (default: 1 0)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sisi4s_b200 import synthetic as S
from sisi4s_b200.triples import TriplesEngine

orders = [int(a) for a in sys.argv[1:]] or [1, 0]
out = {}
inp = S.make_inputs(20, 100, seed=2026, kind="vertex")
with TriplesEngine(20, 100) as eng:
    eng.set_inputs(*inp.args())
    for order in orders:
        eng.set_option("order", order)
        for rep in range(3):
            r = eng.run()
        out[f"o20_v100_order{order}"] = {"E": r.energy, "s_kernel": r.seconds_kernel, "tflops": r.flops / r.seconds_kernel * 1e-12}
        print(json.dumps(out[f"o20_v100_order{order}"]), flush=True)
t0 = time.time()
inp = S.make_inputs(40, 300, seed=2026, kind="vertex", nf=24)
print("gen s", time.time() - t0, flush=True)
with TriplesEngine(40, 300) as eng:
    t0 = time.time(); eng.set_inputs(*inp.args()); st = eng.stats()
    print("upload+pack wall", time.time() - t0, "dev", st.seconds_upload, "GB", st.bytes_h2d / 1e9, flush=True)
    for order in orders:
        eng.set_option("order", order)
        for part in (3, 3, 6):
            b, e = eng.partition(8, part)
            r = eng.run(b, e)
            key = f"o40_v300_order{order}_part{part}"
            out[key] = {"E": r.energy, "s_kernel": r.seconds_kernel, "tflops": r.flops / r.seconds_kernel * 1e-12}
            print(key, json.dumps(out[key]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/quick_timing.json", "w"), indent=1)
