"""BASELINE.json configs[3] shape on ONE GPU: synthetic (T) at o=64, v=512 with the whole packed
input set (87 GB) resident.  The PPPH integrals are built on the device from the vertex
(CoulombVertex contract), a weight-balanced 1/NPART of the sorted triples is timed, and sampled
triples are checked against the CPU oracle (oracle/pt_oracle.c), which reads only the three PPPH
slabs of a triple from a sparse memory-mapped [v,v,v,o] file.

    python tests/manual_config4_check.py [o v npart part]     -> gpurun_out/config4_check.json
"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sisi4s_b200 import synthetic as S
from sisi4s_b200.triples import TriplesEngine
from oracle import c_oracle as CO

o, v, npart, part = (int(x) for x in (sys.argv[1:5] + ["64", "512", "64", "20"][len(sys.argv) - 1:]))
out = {"o": o, "v": v}
t0 = time.time()
inp = S.make_inputs(o, v, seed=2026, kind="vertex", nf=24, with_ppph=False)
out["host_generation_s"] = time.time() - t0
tr = [(i, j, k) for i in range(o) for j in range(i, o) for k in range(j, o)]
with TriplesEngine(o, v) as eng:
    t0 = time.time()
    eng.set_eigenenergies(inp.epsi, inp.epsa); eng.set_singles(inp.T1); eng.set_doubles(inp.T2)
    eng.set_pphh(inp.Vpphh); eng.set_hhhp(inp.Vhhhp); eng.set_vertex(inp.Gamma)
    st = eng.stats()
    out.update(setup_wall_s=time.time() - t0, setup_device_s=st.seconds_upload, device_gb=st.device_bytes / 1e9,
               h2d_gb=st.bytes_h2d / 1e9)
    print(json.dumps(out), flush=True)
    b, e = eng.partition(npart, part)
    eng.run(b, b + 8)                       # warm-up
    r = eng.run(b, e)
    out["timed"] = {"triples": [b, e], "fraction": f"1/{npart}", "s_kernel": r.seconds_kernel, "s_run": r.seconds,
                    "tflops": r.flops / r.seconds_kernel * 1e-12, "energy": r.energy,
                    "full_problem_est_s": r.seconds * npart}
    print(json.dumps(out["timed"]), flush=True)
    picks = [tr.index(t) for t in ((3, 17, 40), (5, 5, 30), (7, 21, 21))]
    fused = [eng.run(t, t + 1).energy for t in picks]
# CPU oracle on the sampled triples: sparse PPPH file, only the touched slabs are written
path = os.path.join(tempfile.gettempdir(), f"ppph_{o}_{v}.sparse")
mm = np.memmap(path, dtype=np.float64, mode="w+", shape=(v, v, v, o), order="F")
for z in sorted({h for t in picks for h in tr[t]}):
    mm[:, :, :, z] = S.ppph_slab_from_vertex(inp.Gamma, o, v, z)
t0 = time.time()
ref = CO.triples_list(inp.epsi, inp.epsa, inp.T1, inp.T2, inp.Vpphh, inp.Vhhhp, mm, np.array(picks))
out["oracle"] = {"triples": [tr[t] for t in picks], "fused": fused, "cpu_oracle": [float(x) for x in ref],
                 "max_abs_diff": float(np.abs(np.array(fused) - ref).max()), "cpu_s": time.time() - t0,
                 "cpu_threads": CO.max_threads()}
del mm
os.remove(path)
print(json.dumps(out["oracle"]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/config4_check_o{o}_v{v}.json", "w"), indent=1)
assert out["oracle"]["max_abs_diff"] <= 1e-9
