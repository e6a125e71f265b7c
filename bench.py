#!/usr/bin/env python3
"""Benchmark of the (T) hot path (BASELINE.json metric: FP64 TFLOP/s and wall-s at
o=40, v=300 on 1/2/4/8 B200, next to the CPU path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload: synthetic UEG-style closed-shell inputs at o=40, v=300 (BASELINE configs[2]),
11 GB of integrals/amplitudes per GPU (replicated), far larger than the 126 MB L2.
One STEP = one weight-balanced contiguous eighth of the sorted-triple list
(i<=j<=k, reference enumeration order) processed by one `pt_run` per rank; eight
consecutive steps = one complete E(T).  With N ranks every step's eighth is split
N ways (pt_partition), so the total work is fixed: strong scaling.  The only
communication is one all-reduce of the scalar energy (and of the timings).

`value` = algorithmic FLOP (2 v^3 (v+o) per ordered hole triple) of the timed steps
of all ranks / max-over-ranks device time, inputs already resident in HBM.
`e2e` = the same metric through the plugin-level API (sisi4s_b200.triples.
CcsdPerturbativeTriples.run): host buffers -> H2D copies + packing + all triples
+ D2H of the energy, timed on the host around the call.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# Host threads.  torchrun exports OMP_NUM_THREADS=1 unless it is already set; the (untimed) input
# generation then shares the host cores between the N ranks, and the reference arm -- which runs on
# rank 0 alone -- takes all of them.
if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
    _cores = os.cpu_count() or 8
    _ref = "--impl" in sys.argv and "reference" in sys.argv
    os.environ["OMP_NUM_THREADS"] = str(_cores if _ref else max(1, _cores // int(os.environ["WORLD_SIZE"])))

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

O_, V_ = 40, 300
NBATCH = 8
NF_SYNTH = 24  # auxiliary index of the synthetic vertex (setup cost only; not on the timed path)
# --workload: (o, v, steps per complete E(T), PPPH on the host?, BASELINE.json config)
WORKLOADS = {
    "o40v300": (40, 300, 8, True, "configs[2]"),     # the configuration the metric is quoted on (default)
    "o20v100": (20, 100, 1, True, "configs[1]"),
    "o64v512": (64, 512, 64, False, "configs[3]"),   # 96.6 GB per GPU; PPPH built on the device from the vertex
}
HOST_PPPH = True
CONFIG_NAME = "configs[2]"


def flops_of(o, v, weight):
    return 2.0 * v ** 3 * (v + o) * weight


def triple_weights(o):
    return np.array([[6, 3, 3, 1][(i == j) + 2 * (j == k)]
                     for i in range(o) for j in range(i, o) for k in range(j, o)], dtype=np.int64)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measure_fp64_peak(device):
    """cuBLAS DGEMM 8192^3 through torch (library GEMM as the measured FP64 ceiling;
    MEASURED_PEAKS.json carries no FP64 entry)."""
    import torch
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    torch.matmul(a, b)
    torch.cuda.synchronize(device)
    best = 0.0
    t_end = time.time() + 4.0
    rates = []
    while time.time() < t_end:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); e1.synchronize()
        rates.append(2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) * 1e-12)
        best = max(best, rates[-1])
    del a, b
    torch.cuda.empty_cache()
    return best, float(np.median(rates[len(rates) // 2:]))


def pinned_like(arr):
    """Copy a numpy array into page-locked host memory, same shape / Fortran order."""
    import torch
    t = torch.empty(arr.size, dtype=torch.float64, pin_memory=True)
    out = t.numpy().reshape(arr.shape, order="F")
    out[...] = arr
    return out, t


def cpu_baseline_sample(inp, idx_pool, weights, budget_s=12.0, max_s=30.0):
    """C restatement of the reference loop (oracle/pt_oracle.c) on all host cores, one sorted
    triple at a time from idx_pool until the time budget is spent."""
    from oracle import c_oracle as CO
    t0 = time.time()
    done, w = [], 0
    for t in idx_pool:
        CO.triples_list(*inp.args(), np.array([t]))
        done.append(int(t)); w += int(weights[t])
        el = time.time() - t0
        if el >= budget_s or el + el / len(done) > max_s:
            break
    el = time.time() - t0
    return {"value": flops_of(inp.o, inp.v, w) / el * 1e-12, "unit": "TFLOP/s", "cores": CO.max_threads(),
            "kind": "port", "seconds": el,
            "sample": f"{len(done)} sorted triples {done} of the o={inp.o},v={inp.v} workload "
                      f"({w} W blocks, {flops_of(inp.o, inp.v, w):.3e} FLOP)"}


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU algorithm for the path.  sisi4s itself (MPI +
    Cyclops CTF) cannot be built here, so this times the C restatement of its loop
    (oracle/pt_oracle.c, OpenMP over all host cores) on a bounded sample per step."""
    if rank != 0:
        return
    if not HOST_PPPH:
        print(json.dumps({"impl": "reference", "unavailable": f"workload {args.workload}: the CPU arm needs the "
                          "v^3 o PPPH tensor on the host; timed at o40v300 / o20v100 only"}), flush=True)
        return
    from oracle import c_oracle as CO
    from sisi4s_b200 import synthetic as S
    inp = S.make_inputs(O_, V_, seed=2026, kind="vertex", nf=NF_SYNTH)
    w = triple_weights(O_)
    ntr = w.size
    per_step = 2  # sorted triples per step (bounded sample of that step's eighth)
    def step_triples(s):
        b = (s % NBATCH) * ntr // NBATCH
        return np.array([b + 17 + 3 * q for q in range(per_step)])
    for s in range(args.warmup):
        CO.triples_list(*inp.args(), step_triples(s)[:1])
    t0 = time.time()
    fl = 0.0
    for s in range(args.steps):
        idx = step_triples(s)
        CO.triples_list(*inp.args(), idx)
        fl += flops_of(O_, V_, int(w[idx].sum()))
    el = time.time() - t0
    val = fl / el * 1e-12
    line = {
        "impl": "reference", "metric": f"(T) FP64 TFLOP/s at o={O_},v={V_}", "value": val, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic UEG-style (T), o={O_} v={V_} (BASELINE {CONFIG_NAME})", "o": O_, "v": V_,
                   "step": f"bounded sample: {per_step} sorted triples of the step's eighth of the triple list"},
        "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": CO.max_threads(), "kind": "port",
                         "sample": f"{per_step} sorted triples per step, {args.steps} steps; C restatement of "
                                   "CcsdPerturbativeTriples.cxx:159-216 (sisi4s needs MPI+CTF, unbuildable here)"},
        "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="o40v300", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    global O_, V_, NBATCH, HOST_PPPH, CONFIG_NAME
    O_, V_, NBATCH, HOST_PPPH, CONFIG_NAME = WORKLOADS[args.workload]
    if args.steps is None:
        args.steps = min(NBATCH, 8)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from sisi4s_b200 import synthetic as S
    from sisi4s_b200.triples import TriplesEngine, CcsdPerturbativeTriples
    from sisi4s_b200.sharding import TripleShards

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the (T) path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL announces its version on stdout when the first communicator is created; keep stdout
        # for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    shards = TripleShards(O_, world, rank)

    def allreduce(x, op):
        return shards.max(x, dev) if op is dist.ReduceOp.MAX else shards.sum(x, dev)

    # ---- setup (untimed): inputs, FP64 ceiling, upload + pack
    t_setup = time.time()
    inp = S.make_inputs(O_, V_, seed=2026, kind="vertex", nf=NF_SYNTH, with_ppph=HOST_PPPH)
    # the large tensors live in page-locked host memory from here on (one copy per rank: the
    # pageable originals are dropped, so that 8 ranks fit the host's memory)
    keep = []
    for field in ("T2", "Vpphh", "Vppph", "Vhhhp") if HOST_PPPH else ("T2", "Vpphh", "Vhhhp"):
        view, owner = pinned_like(getattr(inp, field))
        setattr(inp, field, view)
        keep.append(owner)
    weights = triple_weights(O_)
    peak_burst, peak_sust = measure_fp64_peak(dev)
    eng = TriplesEngine(O_, V_, device=local)
    eng.set_inputs(inp.epsi, inp.epsa, inp.T1, inp.T2, inp.Vpphh, inp.Vhhhp, inp.Vppph,
                   vertex=None if HOST_PPPH else inp.Gamma)
    t_setup = time.time() - t_setup

    def step_range(s):
        return shards.my_range(NBATCH, s)

    for s in range(args.warmup):
        eng.run(*step_range(s))
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_s = ker_s = fl = 0.0
    e_sum = 0.0
    launches0 = eng.stats().kernel_launches
    wall0 = time.time()
    for s in range(args.steps):
        b, e = step_range(s)
        res = eng.run(b, e)
        dev_s += res.seconds; ker_s += res.seconds_kernel; fl += res.flops; e_sum += res.energy
    barrier()
    wall = time.time() - wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.stats().kernel_launches - launches0
    t_max = allreduce(dev_s, dist.ReduceOp.MAX)
    k_max = allreduce(ker_s, dist.ReduceOp.MAX)
    fl_all = allreduce(fl, dist.ReduceOp.SUM)
    e_all = allreduce(e_sum, dist.ReduceOp.SUM)   # the one data collective
    launches_all = int(allreduce(float(launches), dist.ReduceOp.SUM))
    value = fl_all / t_max * 1e-12
    eng.close()

    # ---- end-to-end leg: plugin API, host (pinned) buffers -> energy
    e2e = None
    if not args.no_e2e and NBATCH <= 8:   # larger workloads: one complete E(T) takes > 20 min on one GPU
        big = dict(CcsdDoublesAmplitudes=inp.T2, PPHHCoulombIntegrals=inp.Vpphh,
                   HHHPCoulombIntegrals=inp.Vhhhp)   # pinned (see setup)
        if HOST_PPPH:
            big["PPPHCoulombIntegrals"] = inp.Vppph
        else:
            big["CoulombVertex"] = inp.Gamma         # PPPH is built on the device
        data = dict(HoleEigenEnergies=inp.epsi, ParticleEigenEnergies=inp.epsa, CcsdEnergy=inp.ccsd_energy,
                    CcsdSinglesAmplitudes=inp.T1, **big)
        argsmap = {k: "$" + k for k in data}
        argsmap["PerturbativeTriplesEnergy"] = "$PerturbativeTriplesEnergy"
        argsmap["device"] = local

        class Shard(CcsdPerturbativeTriples):
            """the plugin run() restricted to this rank's share of the sorted triples"""
            def run(self):
                with self.make_engine() as en:
                    r = en.run(*shards.my_range())
                    self.stats = en.stats()
                return r

        barrier()
        w0 = time.time()
        alg = Shard(argsmap, data)
        r = alg.run()
        e_e2e = allreduce(r.energy, dist.ReduceOp.SUM)
        barrier()
        w_e2e = allreduce(time.time() - w0, dist.ReduceOp.MAX)
        fl_e2e = flops_of(O_, V_, int(weights.sum()))
        e2e = {"value": fl_e2e / w_e2e * 1e-12, "unit": "TFLOP/s", "seconds": w_e2e,
               "h2d_bytes_per_step": float(alg.stats.bytes_h2d), "d2h_bytes_per_step": float(alg.stats.bytes_d2h),
               "step": "one complete E(T): upload + pack + all sorted triples + energy read-back",
               "energy": e_e2e + inp.ccsd_energy, "triples_energy": e_e2e}

    cpu = None
    if rank == 0 and world == 1 and HOST_PPPH and not args.no_cpu:   # reported at N=1 only
        pool = [int(t) for t in np.linspace(40, weights.size - 40, 24).astype(int)]
        cpu = cpu_baseline_sample(inp, pool, weights)

    if rank == 0:
        traffic = None
        prof = os.path.join(ROOT, "profiles", "fused_kernel_traffic.json")
        if os.path.exists(prof) and args.workload == "o40v300":   # the capture is of this workload's step
            try:
                # measured on one N=1 step (profiles/r01d_step_traffic.csv); a rank's launch covers 1/world of it
                traffic = json.load(open(prof)).get("dram_bytes_per_launch") / world
            except Exception:
                traffic = None
        achieved = fl_all / world / k_max * 1e-12   # per GPU: the roofline is the kernel's, not the job's
        line = {
            "metric": f"(T) FP64 TFLOP/s at o={O_},v={V_}", "value": value, "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_max / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"synthetic UEG-style (T), o={O_} v={V_} (BASELINE {CONFIG_NAME})", "o": O_, "v": V_,
                       "step": f"1/{NBATCH} of the sorted-triple list (weight-balanced contiguous chunk) per step, "
                               f"split over {world} rank(s); {NBATCH} steps = one complete E(T)",
                       "parallelism": f"triples sharded over {world} GPU(s), inputs replicated",
                       "l2": "inputs (>= 0.2 GB/GPU packed; 11 GB at o=40,v=300) exceed the 126 MB L2; no flush needed",
                       "triples_energy_of_timed_steps": e_all, "wall_s_timed": wall,
                       "wall_s_full_problem_est": t_max / args.steps * NBATCH, "setup_s": t_setup},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_burst, "unit": "TFLOP/s",
                         "frac": achieved / peak_burst, "traffic": traffic,
                         "peak_source": "cuBLAS DGEMM 8192^3 via torch.matmul measured in this run (burst; "
                                        f"sustained median {peak_sust:.2f}); MEASURED_PEAKS.json has no FP64 entry",
                         "frac_of_nominal_37": achieved / 37.0,
                         "kernel": "pt_fused_kernel", "algorithmic_flop": fl_all / world,
                         "per": "GPU (slowest rank's kernel time)"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_all, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
