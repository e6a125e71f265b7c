"""Committed golden fixtures that are too expensive to regenerate inside the test run."""
import numpy as np


def test_o40v300_triples_fixture_is_well_formed():
    """tests/golden/o40v300_triples.json (make_o40v300_triples.py): 200 distinct sorted triples of the metric's
    shape in the reference's enumeration (CcsdPerturbativeTriples.cxx:156-158), all four hole classes present."""
    import json
    import os
    from oracle import c_oracle as CO
    with open(os.path.join(os.path.dirname(__file__), "golden", "o40v300_triples.json")) as f:
        g = json.load(f)
    e = np.array([float.fromhex(x) for x in g["energy"]])
    assert len(set(g["index"])) == len(g["index"]) == e.size >= 200
    assert all(tuple(t) == CO.triple_of(g["o"], n) for n, t in zip(g["index"], g["triple"]))
    classes = {(i == j, j == k) for i, j, k in g["triple"]}
    assert classes == {(False, False), (True, False), (False, True), (True, True)}
    assert np.all(np.isfinite(e)) and np.all(e <= 0.0) and abs(e.sum() - g["sum"]) < 1e-18
