#!/usr/bin/env python3
"""FP64 ceilings of the device: DMMA.8x8x4 and DFMA issue rates (pt_bench_fp64) and the
cuBLAS DGEMM rate (torch.matmul), printed as one JSON object.  Run on the B200 box."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from sisi4s_b200.triples import TriplesEngine  # noqa: E402

out = {"device": torch.cuda.get_device_name(0), "sm_count": torch.cuda.get_device_properties(0).multi_processor_count}
with TriplesEngine(2, 4) as eng:
    for mode, name in ((0, "dmma"), (1, "dfma")):
        for warps in (4, 8, 16, 32):
            tf, mhz = eng.bench_fp64(mode, warps, 20000)
            out[f"{name}_w{warps}"] = {"tflops": tf, "sm_mhz_est": mhz}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b); torch.cuda.synchronize()
    best, rates = 0.0, []
    t_end = time.time() + 3.0
    while time.time() < t_end:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); e1.synchronize()
        rates.append(2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) * 1e-12)
    out[f"dgemm_{n}"] = {"burst": max(rates), "sustained_median": sorted(rates[len(rates) // 2:])[len(rates) // 4]}
# the (T) GEMM shape: (v x v) . (v x v^2) at v = 300
v = 300
a = torch.randn(v, v, dtype=torch.float64, device="cuda")
b = torch.randn(v, v * v, dtype=torch.float64, device="cuda")
torch.matmul(a, b); torch.cuda.synchronize()
rates = []
for _ in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); e1.synchronize()
    rates.append(2.0 * v ** 4 / (e0.elapsed_time(e1) * 1e-3) * 1e-12)
out["dgemm_300x300x90000"] = {"burst": max(rates), "median": sorted(rates)[len(rates) // 2]}
print(json.dumps(out, indent=1))
