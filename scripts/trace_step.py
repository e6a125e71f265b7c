"""Where a (40,300) bench step spends its time inside pt_run (PT_TRACE=1 prints host-side marks)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PT_TRACE"] = "1"
import torch
from sisi4s_b200 import synthetic as S
from sisi4s_b200.triples import TriplesEngine
import bench
dev = torch.device("cuda", 0)
host = bench.HostBuffers(False, 0, lambda: None, "trace")
inp = bench.generate_inputs(bench.WORKLOADS["o40v300"], dev, host, 0, 1)
with TriplesEngine(40, 300) as eng:
    eng.set_inputs(*inp.args())
    for part in (0, 1):
        b, e = eng.partition(64, part)
        r = eng.run(b, e)
        print("part", part, "s_run", r.seconds, "s_kernel", r.seconds_kernel, flush=True)
    b, e = eng.partition(8, 3)
    r = eng.run(b, e)
    print("step", "s_run", r.seconds, "s_kernel", r.seconds_kernel, flush=True)
