"""A/B timing of library builds and debug switches on one (40,300) bench step.
usage: python scripts/ab_timing.py lib1.so[:debug,debug,...] lib2.so[:...]   (run each in a subprocess)"""
import json, os, subprocess, sys
HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys
sys.path.insert(0, %r)
import torch
import bench
from sisi4s_b200.triples import TriplesEngine
dbg = [x for x in sys.argv[1].split(",")]
inp = bench.generate_inputs(bench.WORKLOADS["o40v300"], torch.device("cuda", 0), bench.HostBuffers(False, 0, lambda: None, "ab"), 0, 1)
with TriplesEngine(40, 300) as eng:
    eng.set_inputs(*inp.args())
    b, e = eng.partition(8, 3)
    eng.run(b, b + 200)
    for d in dbg:
        # "N" = debug switch N; "key=value" = any other option
        if "=" in d:
            k, val = d.split("="); eng.set_option(k, int(val))
        else:
            eng.set_option("debug", int(d))
        r = eng.run(b, e)
        print(json.dumps({"lib": os.path.basename(os.environ.get("SISI4S_PT_LIB", "default")), "debug": d,
                          "s_kernel": r.seconds_kernel, "tflops_equiv": r.flops / r.seconds_kernel * 1e-12,
                          "E": r.energy}), flush=True)
''' % HERE
for spec in sys.argv[1:]:
    lib, _, dbg = spec.partition(":")
    env = dict(os.environ)
    if lib != "default":
        env["SISI4S_PT_LIB"] = os.path.join(HERE, "sisi4s_b200", lib)
    subprocess.run([sys.executable, "-c", CHILD, dbg or "0"], env=env, check=False)
