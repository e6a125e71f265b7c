#!/usr/bin/env python3
"""Summarise an .ncu-rep of pt_fused_kernel: headline counters, warp-sample shares per code
region (regions = runs of SASS instructions with equal execution counts) and the top stall
reasons.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [min_share]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
for k in ["gpu__time_duration.sum", "sm__cycles_elapsed.max",
          "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
          "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
          "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "smsp__inst_executed.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
          "smsp__issue_active.avg.pct_of_peak_sustained_active"]:
    print(f"{k} = {d.get(k)} {u.get(k)}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
tot = sum(int(r[ix["# Samples"]]) for r in data)
S = lambda a, b: sum(int(r[ix["# Samples"]]) for r in data[a:b])
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total warp samples", tot)
prev = None; start = 0
for n, r in enumerate(data + [None]):
    c = r[ix["Instructions Executed"]] if r else None
    if c != prev:
        if prev is not None and S(start, n) > tot * min_share:
            agg = {s[6:]: sum(int(q[ix[s]] or 0) for q in data[start:n]) for s in stalls}
            top = sorted(agg.items(), key=lambda x: -x[1])[:4]
            kinds = [k for k in ("DMMA", "LDTM", "STTM", "UBLKCP", "BAR.SYNC", "SYNCS", "LDS", "STS", "LDG", "MUFU", "ATOM")
                     if any(k in q[ix["Source"]] for q in data[start:n])]
            print(f"[{start},{n}) x{prev} {100*S(start,n)/tot:5.1f}%  {top}  {kinds}")
        start = n; prev = c
agg = {s[6:]: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print("stalls:", sorted(agg.items(), key=lambda x: -x[1])[:8])
