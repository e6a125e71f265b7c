"""Per-triple (T) energies of 200 seeded sorted triples at BASELINE configs[2] (o=40, v=300), computed on the
CPU by the C restatement of the reference (oracle/pt_oracle.c through oracle/c_oracle.py, GEMMs in OpenBLAS)
on the host-generated synthetic inputs (synthetic.make_inputs(seed=2026, kind="vertex", nf=24)).

    python tests/golden/make_o40v300_triples.py        # ~10 min on 8 cores, writes o40v300_triples.json

The sample is the >=200-triple sample BASELINE.md section 3 asks for: 184 triples drawn uniformly (seed 2026)
plus 16 hand-picked ones that cover the four hole classes (i<j<k, i=j<k, i<j=k, i=j=k) and the corners.
tests/test_gpu_round2.py compares the CUDA path with it on the GPU box, where the oracle has time for ~26.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import c_oracle as CO          # noqa: E402
from sisi4s_b200 import synthetic as S     # noqa: E402

O, V, SEED, NF = 40, 300, 2026, 24
HAND = ((0, 1, 2), (3, 17, 39), (20, 21, 22), (11, 30, 38), (37, 38, 39), (0, 19, 39), (0, 0, 1), (12, 12, 30),
        (38, 38, 39), (0, 1, 1), (7, 25, 25), (0, 39, 39), (39, 39, 39), (5, 5, 5), (0, 0, 0), (0, 0, 39))


def main():
    triples = [(i, j, k) for i in range(O) for j in range(i, O) for k in range(j, O)]
    hand = [triples.index(t) for t in HAND]
    rng = np.random.default_rng(SEED)
    drawn = [int(t) for t in rng.permutation(len(triples)) if int(t) not in hand][:200 - len(hand)]
    idx = np.array(hand + drawn, dtype=np.int64)
    t0 = time.time()
    inp = S.make_inputs(O, V, seed=SEED, kind="vertex", nf=NF)
    print(f"inputs {time.time() - t0:.1f} s", flush=True)
    CO.use_blas(True)
    t0 = time.time()
    e = CO.triples_list(*inp.args(), idx)
    print(f"oracle: {idx.size} triples in {time.time() - t0:.1f} s", flush=True)
    out = {"_comment": "E_t of 200 sorted triples at o=40, v=300 (reference enumeration index -> energy), C oracle on "
                       "synthetic.make_inputs(seed=2026, kind='vertex', nf=24); made by make_o40v300_triples.py",
           "o": O, "v": V, "seed": SEED, "nf": NF, "index": [int(t) for t in idx],
           "triple": [list(triples[int(t)]) for t in idx], "energy": [float(x).hex() for x in e],
           "sum": float(np.sum(e))}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "o40v300_triples.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("sum", out["sum"])


if __name__ == "__main__":
    main()
