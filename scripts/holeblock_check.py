"""In-library hole-block mode at (40,300), block 8 (>= 0.85 of the resident throughput
there; the Python driver of round 1 reached 0.49): one bench step (partition 3 of 8) through
pt_set_option("hole_block", 8) with T2 / PPHH / PPPH in host memory, against the all-resident engine --
throughput including the per-group staging, and bitwise equality of the per-triple energies.
    python scripts/holeblock_check.py -> gpurun_out/holeblock_check_o40_v300.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from sisi4s_b200.triples import TriplesEngine
dev = torch.device("cuda", 0)
host = bench.HostBuffers(False, 0, lambda: None, "hb")
inp = bench.generate_inputs(bench.WORKLOADS["o40v300"], dev, host, 0, 1)
out = {"o": 40, "v": 300, "block": 8}
with TriplesEngine(40, 300) as eng:
    eng.set_inputs(*inp.args())
    b, e = eng.partition(8, 3)
    eng.run(b, b + 100)
    r = eng.run(b, e)
    base = r.per_triple
    out["resident"] = {"s_run": r.seconds, "s_kernel": r.seconds_kernel, "tflops": r.flops / r.seconds * 1e-12}
for source in ("host_ppph", "vertex", "vertex_only"):
    with TriplesEngine(40, 300, hole_block=8) as eng:
        if source == "host_ppph":
            eng.set_inputs(inp.epsi, inp.epsa, inp.T1, inp.T2, inp.Vpphh, inp.Vhhhp, inp.Vppph)
        elif source == "vertex":
            eng.set_inputs(inp.epsi, inp.epsa, inp.T1, inp.T2, inp.Vpphh, inp.Vhhhp, vertex=inp.Gamma)
        else:
            eng.set_inputs(inp.epsi, inp.epsa, inp.T1, inp.T2, None, None, vertex=inp.Gamma)
        eng.run(b, b + 100)
        st0 = eng.stats()
        r = eng.run(b, e)
        st = eng.stats()
        diff = float(np.abs(r.per_triple - base).max())
        out[source] = {"s_run": r.seconds, "s_kernel": r.seconds_kernel, "tflops_incl_staging": r.flops / r.seconds * 1e-12,
                       "fraction_of_resident": out["resident"]["s_run"] / r.seconds, "groups_staged": int(st.groups_staged - st0.groups_staged),
                       "slab_loads": int(st.slab_loads - st0.slab_loads), "bitwise_equal_to_resident": bool(np.array_equal(r.per_triple, base)),
                       "max_abs_diff": diff, "device_gb": st.device_bytes / 1e9}
    print(source, json.dumps(out[source]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/holeblock_check_o40_v300.json", "w"), indent=1)
