"""CUDA path vs the committed 200-triple CPU fixture at o=40, v=300 (tests/golden/o40v300_triples.json):
prints the largest per-triple and summed differences as one JSON line (profiles/r02h_parity_200.json)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sisi4s_b200 import synthetic as S                  # noqa: E402
from sisi4s_b200.triples import TriplesEngine           # noqa: E402

with open(os.path.join(ROOT, "tests", "golden", "o40v300_triples.json")) as f:
    g = json.load(f)
want = np.array([float.fromhex(x) for x in g["energy"]])
inp = S.make_inputs(g["o"], g["v"], seed=g["seed"], kind="vertex", nf=g["nf"])
with TriplesEngine(g["o"], g["v"]) as eng:
    eng.set_inputs(inp.epsi, inp.epsa, inp.T1, inp.T2, inp.Vpphh, inp.Vhhhp, inp.Vppph)
    got = eng.run_list(g["index"]).per_triple
d = np.abs(got - want)
print(json.dumps({"workload": "o40v300", "triples": int(want.size), "max_abs_diff": float(d.max()),
                  "max_rel_diff": float((d / np.abs(want))[np.abs(want) > 1e-12].max()),   # i=j=k triples are exactly 0 up to rounding
                  "sum_diff": float(abs(got.sum() - want.sum())),
                  "sum": float(got.sum()), "fixture": "tests/golden/o40v300_triples.json (CPU: oracle/pt_oracle.c + OpenBLAS)"}))
