"""One packed PPPH slab + one PPHH block from the vertex (vertex_gemm_kernel) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sisi4s_b200 import synthetic as S
from sisi4s_b200.triples import TriplesEngine
o, v, nf = (int(x) for x in (sys.argv[1:4] + ["24", "512", "1024"][len(sys.argv) - 1:]))
G = S.make_vertex(o, v, seed=7, nf=nf)
with TriplesEngine(o, v, slab_slots=3) as eng:
    eng.set_vertex(G)
    s, f = eng.bench_vertex_gemm(0, 1)
    print("slab", s, f / s * 1e-12)
