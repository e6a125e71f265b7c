"""Integrals-from-vertex GEMM (pt_pack.cu: vertex_gemm_kernel) at realistic auxiliary dimensions:
kernel-only device time of one packed PPPH slab and of the PPHH block, as FP64 TFLOP/s.

    python scripts/vertex_timing.py [o v nf ...]   -> gpurun_out/vertex_timing.json
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sisi4s_b200 import synthetic as S
from sisi4s_b200.triples import TriplesEngine
a = [int(x) for x in sys.argv[1:]] or [40, 300, 600, 40, 300, 24, 24, 512, 1024]
out = []
for o, v, nf in zip(a[0::3], a[1::3], a[2::3]):
    G = S.make_vertex(o, v, seed=7, nf=nf)
    with TriplesEngine(o, v, slab_slots=3 if o > 3 else 0) as eng:   # 3 slots: only the vertex + 3 slabs live
        eng.set_vertex(G)
        rec = {"o": o, "v": v, "nf": nf}
        for what, name in ((0, "ppph_slab_packed"), (1, "pphh")):
            s, f = eng.bench_vertex_gemm(what, 3)
            rec[name] = {"seconds": s, "flop": f, "tflops": f / s * 1e-12, "out_gb_per_s": (v ** 3 if what == 0 else v * v * o * o) * 8 / s * 1e-9}
        out.append(rec)
        print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/vertex_timing.json", "w"), indent=1)
