"""Device construction of the PPPH slabs from the Coulomb vertex (pt_set_vertex) at a realistic
auxiliary dimension: time + one slab checked against numpy through the W-tile debug entry."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sisi4s_b200 import synthetic as S
from sisi4s_b200.triples import TriplesEngine
o, v, nf = (int(x) for x in (sys.argv[1:4] + ["40", "300", "600"][len(sys.argv) - 1:]))
G = S.make_vertex(o, v, seed=7, nf=nf)
inp = S.make_inputs(o, v, seed=7, kind="random") if v <= 64 else None
with TriplesEngine(o, v) as eng:
    t0 = time.time()
    eng.set_vertex(G)
    st = eng.stats()
    wall = time.time() - t0
    flop = 2.0 * 2.0 * nf * v ** 3 * o
    h2d_s = st.bytes_h2d / 25e9
    out = {"o": o, "v": v, "nf": nf, "device_s_total": st.seconds_upload, "wall_s": wall, "flop": flop,
           "tflops_incl_h2d_and_pack": flop / st.seconds_upload * 1e-12, "h2d_gb": st.bytes_h2d / 1e9}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/vertex_timing.json", "w"), indent=1)
