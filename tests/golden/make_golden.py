#!/usr/bin/env python3
"""Regenerates tests/golden/triples_golden.npz with the NumPy oracle.

The reference (sisi4s) cannot be built or imported here (needs MPI + Cyclops
CTF), so these are outputs of oracle/pt_oracle.py form A (the literal restatement
of CcsdPerturbativeTriples.cxx:119-248), cross-checked against form B
(PerturbativeTriples.cxx:172-239) wherever form B fits in memory.  Inputs are
regenerated bit-identically from sisi4s_b200.synthetic by (o, v, kind, seed).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pt_oracle as O  # noqa: E402
from sisi4s_b200 import synthetic as S  # noqa: E402

CASES = [  # (o, v, kind, seed)
    (1, 3, "random", 4), (2, 5, "random", 4), (3, 16, "random", 17), (3, 17, "random", 17),
    (5, 19, "random", 2026), (5, 19, "vertex", 2026), (4, 33, "vertex", 17), (6, 40, "random", 4),
    (20, 100, "vertex", 2026),
]

out = {}
for (o, v, kind, seed) in CASES:
    t0 = time.time()
    inp = S.make_inputs(o, v, seed=seed, kind=kind)
    e, per = O.triples_loop(*inp.args(), return_per_triple=True)
    key = f"o{o}_v{v}_{kind}_s{seed}"
    if o ** 3 * v ** 3 <= 3e6:
        eb = O.triples_full(*inp.args())
        assert abs(e - eb) <= 1e-12 * max(1.0, abs(e)), (key, e, eb)
    out[key + "_total"] = np.array(e)
    out[key + "_per_triple"] = per
    print(f"{key}: E(T) = {e:.15e}   ({time.time() - t0:.1f} s)", flush=True)
np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "triples_golden.npz"), **out)
