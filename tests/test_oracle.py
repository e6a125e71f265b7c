"""CPU tests of the oracle (test infrastructure) against itself and the golden fixtures."""
import os

import numpy as np
import pytest

from oracle import pt_oracle as O
from sisi4s_b200 import synthetic as S

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "triples_golden.npz"))


def test_permutation_table_matches_survey():
    # Permutation<3>(p).images, reference src/math/Permutation.hpp:52-62
    assert O.PERM == [(0, 1, 2), (1, 0, 2), (1, 2, 0), (0, 2, 1), (2, 0, 1), (2, 1, 0)]
    assert [O.str_after("abc", p) for p in O.PERM] == ["abc", "bac", "bca", "acb", "cab", "cba"]
    sf = [O.SPIN_AND_FERMI[O.invariant_elements_count(p)] for p in O.PERM]
    assert sf == [8.0, -4.0, 2.0, -4.0, 2.0, -4.0]


def test_distinct_permutation_sets():
    def distinct(h):
        return [p for p in range(6)
                if all(O.map_after(h, O.PERM[q]) != O.map_after(h, O.PERM[p]) for q in range(p))]
    assert distinct((0, 1, 2)) == [0, 1, 2, 3, 4, 5]
    assert distinct((0, 0, 1)) == [0, 2, 4]
    assert distinct((0, 1, 1)) == [0, 1, 2]
    assert distinct((1, 1, 1)) == [0]


@pytest.mark.parametrize("o,v,kind", [(3, 4, "random"), (2, 5, "random"), (4, 6, "random"),
                                      (3, 4, "vertex"), (1, 3, "random")])
def test_loop_form_equals_full_form(o, v, kind):
    inp = S.make_inputs(o, v, seed=4, kind=kind, kappa=1.0 if kind == "vertex" else None)
    a = O.triples_loop(*inp.args())
    b = O.triples_full(*inp.args())
    assert abs(a - b) <= 1e-12 * max(1.0, abs(a))


def test_piecuch_form_needs_symmetric_pphh():
    # PerturbativeTriples::runPiecuch (:99-170) equals run (:172-239) when
    # Vabij[a,b,i,j] = Vabij[b,a,j,i]; physical inputs always satisfy this.
    inp = S.make_inputs(3, 5, seed=17, kind="vertex", kappa=1.0)
    assert abs(O.triples_loop(*inp.args()) - O.triples_piecuch(*inp.args())) < 1e-12
    rnd = S.make_inputs(3, 5, seed=17, kind="random")
    rnd.Vpphh = 0.5 * (rnd.Vpphh + rnd.Vpphh.transpose(1, 0, 3, 2))
    assert abs(O.triples_loop(*rnd.args()) - O.triples_piecuch(*rnd.args())) < 1e-11


def test_iii_triples_vanish_identically():
    # sum_sigma sf(sigma) = 8 - 3*4 + 2*2 = 0 makes every i=j=k term vanish (any inputs)
    inp = S.make_inputs(3, 6, seed=2026, kind="random")
    for i in range(3):
        e = O.triple_energy(*inp.args(), (i, i, i))
        assert abs(e) < 1e-10


def test_vertex_integrals_match_reference_formulas():
    inp = S.make_inputs(3, 5, seed=17, kind="vertex")
    V = O.integrals_from_vertex(inp.Gamma, 3, 5)
    for x, y in zip(V, (inp.Vpphh, inp.Vhhhp, inp.Vppph)):
        assert np.abs(x - y).max() < 1e-15
    P = inp.Vpphh
    assert np.abs(P - P.transpose(1, 0, 3, 2)).max() == 0.0


@pytest.mark.parametrize("key", ["o2_v5_random_s4", "o3_v17_random_s17", "o5_v19_vertex_s2026",
                                 "o5_v19_random_s2026", "o4_v33_vertex_s17"])
def test_oracle_reproduces_golden(key):
    o, v, kind, seed = key.split("_")
    inp = S.make_inputs(int(o[1:]), int(v[1:]), seed=int(seed[1:]), kind=kind)
    e, per = O.triples_loop(*inp.args(), return_per_triple=True)
    assert abs(e - float(GOLDEN[key + "_total"])) <= 1e-13 * max(1.0, abs(e))
    assert np.allclose(per, GOLDEN[key + "_per_triple"], rtol=1e-12, atol=1e-15)


def test_generator_is_deterministic():
    a = S.make_inputs(3, 7, seed=2026, kind="vertex")
    b = S.make_inputs(3, 7, seed=2026, kind="vertex")
    assert np.array_equal(a.Vppph, b.Vppph) and np.array_equal(a.T1, b.T1)
    c = S.make_inputs(3, 7, seed=17, kind="vertex")
    assert not np.array_equal(a.Vppph, c.Vppph)
    assert a.Vppph.flags.f_contiguous and a.T2.flags.f_contiguous
