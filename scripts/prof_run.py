"""Short (40,300) run of the fused kernel for ncu captures: one pt_run over a small
range of sorted triples (cold start, inputs resident)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sisi4s_b200.triples import TriplesEngine

# "step P" = the bench's step P (partition P of 8); otherwise <first triple> <count>
step = len(sys.argv) > 1 and sys.argv[1] == "step"
b = 5000 if step or len(sys.argv) <= 1 else int(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
order = int(sys.argv[3]) if len(sys.argv) > 3 else 1
tile_holes = int(sys.argv[4]) if len(sys.argv) > 4 else None
host = bench.HostBuffers(False, 0, lambda: None, "prof")
inp = bench.generate_inputs(bench.WORKLOADS["o40v300"], torch.device("cuda", 0), host, 0, 1)   # the bench's inputs
with TriplesEngine(40, 300) as eng:
    eng.set_inputs(*inp.args())
    eng.set_option("order", order)
    if tile_holes is not None:
        eng.set_option("tile_holes", tile_holes)
    rng = eng.partition(8, n) if step else (b, b + n)
    r = eng.run(*rng)
    print("E", r.energy, "s_kernel", r.seconds_kernel, "TF/s", r.flops / r.seconds_kernel * 1e-12)
