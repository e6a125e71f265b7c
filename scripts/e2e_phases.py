"""Wall-clock phases of the plugin-level (T) run at (40,300): engine creation, setters, pt_run, teardown --
with the lazy wave upload (async setters) and with eager uploads."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sisi4s_b200.triples import TriplesEngine
dev = torch.device("cuda", 0)
host = bench.HostBuffers(False, 0, lambda: None, "e2e")
inp = bench.generate_inputs(bench.WORKLOADS["o40v300"], dev, host, 0, 1)
for mode in ("lazy", "eager", "lazy"):
    t = [time.time()]
    eng = TriplesEngine(40, 300, async_upload=(mode == "lazy")); t.append(time.time())
    eng.set_eigenenergies(inp.epsi, inp.epsa); eng.set_singles(inp.T1); eng.set_doubles(inp.T2)
    eng.set_pphh(inp.Vpphh); eng.set_hhhp(inp.Vhhhp); t.append(time.time())
    (eng.set_ppph_host if mode == "lazy" else eng.set_ppph)(inp.Vppph); t.append(time.time())
    b, e = eng.partition(64, 5)
    r = eng.run(b, e); t.append(time.time())
    st = eng.stats()
    eng.close(); t.append(time.time())
    names = ["create", "setters", "ppph", "run(1/64)", "close"]
    print(mode, {n: round(t[i + 1] - t[i], 3) for i, n in enumerate(names)}, "dev_run", round(r.seconds, 3), "kernel", round(r.seconds_kernel, 3),
          "upload", round(st.seconds_upload, 3), flush=True)
