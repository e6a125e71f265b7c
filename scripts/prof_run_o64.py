"""Short (64,512) run of the fused kernel for an ncu capture: PPPH built on the device from the vertex,
one pt_run over a few sorted triples."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sisi4s_b200.triples import TriplesEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 37
wl = bench.WORKLOADS["o64v512"]
host = bench.HostBuffers(False, 0, lambda: None, "prof64")
inp = bench.generate_inputs(wl, torch.device("cuda", 0), host, 0, 1)
with TriplesEngine(64, 512) as eng:
    eng.set_inputs(inp.epsi, inp.epsa, inp.T1, inp.T2, inp.Vpphh, inp.Vhhhp, vertex=inp.Gamma)
    r = eng.run(20000, 20000 + n)
    print("E", r.energy, "s_kernel", r.seconds_kernel, "TF/s", r.flops / r.seconds_kernel * 1e-12)
