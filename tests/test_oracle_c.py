"""The C restatement of the oracle (oracle/pt_oracle.c, used as CPU baseline and smoke
checker) against the NumPy oracle, and the oracle's permutation algebra against the
reference's own Permutation.hpp compiled from /root/reference (oracle/_ref)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import pt_oracle as O
from sisi4s_b200 import synthetic as S

ORACLE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")


def _c_oracle():
    from oracle import c_oracle as CO
    if not os.path.exists(CO.LIB_PATH):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle_pt.so"])
    return CO


@pytest.mark.parametrize("o,v,kind", [(1, 3, "random"), (3, 17, "random"), (5, 19, "vertex"), (2, 20, "random")])
def test_c_oracle_matches_numpy_oracle(o, v, kind):
    CO = _c_oracle()
    inp = S.make_inputs(o, v, seed=17, kind=kind)
    e, per = O.triples_loop(*inp.args(), return_per_triple=True)
    got = CO.triples_list(*inp.args(), np.arange(per.size))
    assert np.abs(got - per).max() <= 1e-12 * max(1.0, np.abs(per).max())
    # subsets in arbitrary order address the reference enumeration correctly
    idx = np.arange(per.size)[::-2]
    assert np.abs(CO.triples_list(*inp.args(), idx) - per[idx]).max() <= 1e-12 * max(1.0, np.abs(per).max())


def test_permutation_algebra_pinned_to_reference_header():
    """oracle/_ref/permutation_tables.txt is printed by a program that includes the
    reference's src/math/Permutation.hpp (built by `make -C oracle ref` where /root/reference
    exists; the text output travels to the GPU box)."""
    path = os.path.join(ORACLE_DIR, "_ref", "permutation_tables.txt")
    if not os.path.exists(path):
        if os.path.exists("/root/reference/src/math/Permutation.hpp"):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "ref"])
        else:
            pytest.skip("oracle/_ref not built and reference tree absent")
    perms, distinct, compose = {}, {}, {}
    for line in open(path):
        w = line.split()
        if w[0] == "perm":
            perms[int(w[1])] = (tuple(int(x) for x in w[2:5]), int(w[6]), w[8])
        elif w[0] == "distinct":
            distinct[w[1]] = [int(x) for x in w[2:]]
        elif w[0] == "compose":
            compose[(int(w[1]), int(w[2]))] = w[3]
    for p in range(6):
        images, inv, s = perms[p]
        assert images == O.PERM[p]
        assert inv == O.invariant_elements_count(O.PERM[p])
        assert s == O.str_after("abc", O.PERM[p])
    for key, flags in distinct.items():
        h = tuple(int(c) for c in key)
        mine = [int(all(O.map_after(h, O.PERM[q]) != O.map_after(h, O.PERM[p]) for q in range(p)))
                for p in range(6)]
        assert mine == flags
    for (s, p), text in compose.items():
        assert text == O.str_after(O.str_after("abc", O.PERM[s]), O.PERM[p])
