"""Out-of-core hole-block driver (BASELINE configs[4] path): host logic on CPU, parity on the GPU."""
import numpy as np
import pytest

from sisi4s_b200 import synthetic as S
from sisi4s_b200.outofcore import block_groups, deal, run_out_of_core

W = (6, 3, 3, 1)


@pytest.mark.parametrize("o,block", [(7, 2), (7, 3), (10, 4), (5, 8), (100, 8)])
def test_block_groups_tile_the_triple_space(o, block):
    groups = block_groups(o, block)
    ntr = o * (o + 1) * (o + 2) // 6
    seen = sorted(g for _, trip, _ in groups for g, _ in trip)
    assert seen == list(range(ntr))                                   # every sorted triple exactly once
    assert max(len(h) for h, _, _ in groups) <= 3 * block             # active holes per sub-engine
    for holes, trip, w in groups:
        assert all(set(t) <= set(holes) for _, t in trip)
        assert w == sum(W[(i == j) + 2 * (j == k)] for _, (i, j, k) in trip)
    for world in (2, 8):
        parts = [deal(groups, world, r) for r in range(world)]
        assert sum(len(p) for p in parts) == len(groups)
        loads = [sum(g[2] for g in p) for p in parts]
        if len(groups) >= 20 * world:
            assert max(loads) - min(loads) <= 0.05 * max(loads)


@pytest.mark.gpu
@pytest.mark.parametrize("block,source", [(2, "ppph"), (3, "vertex"), (8, "ppph")])
def test_out_of_core_equals_oracle(block, source):
    """Sub-engines over <= 3*block active holes (pt_create_ex / pt_set_doubles_hole / pt_run_list)
    reproduce the per-triple energies of the full problem."""
    from oracle import pt_oracle as O
    o, v = 7, 20
    inp = S.make_inputs(o, v, seed=17, kind="vertex")
    e_ref, per_ref = O.triples_loop(*inp.args(), return_per_triple=True)
    kw = dict(Vppph=inp.Vppph) if source == "ppph" else dict(vertex=inp.Gamma)
    e, per = run_out_of_core(inp.epsi, inp.epsa, inp.T1, inp.T2, inp.Vpphh, inp.Vhhhp, block=block, **kw)
    assert abs(e - e_ref) <= 1e-9
    assert np.abs(per - per_ref).max() <= 1e-9
    parts = [run_out_of_core(inp.epsi, inp.epsa, inp.T1, inp.T2, inp.Vpphh, inp.Vhhhp, block=block,
                             world=2, rank=r, **kw) for r in range(2)]
    assert abs(parts[0][0] + parts[1][0] - e_ref) <= 1e-9
    assert np.abs(parts[0][1] + parts[1][1] - per_ref).max() <= 1e-9


@pytest.mark.gpu
def test_unsymmetric_inputs_through_sub_engines():
    """No input symmetry is assumed by the hole-subset engine either."""
    from oracle import pt_oracle as O
    inp = S.make_inputs(5, 19, seed=2026, kind="random")
    e_ref, per_ref = O.triples_loop(*inp.args(), return_per_triple=True)
    e, per = run_out_of_core(inp.epsi, inp.epsa, inp.T1, inp.T2, inp.Vpphh, inp.Vhhhp, Vppph=inp.Vppph, block=1)
    assert np.allclose(per, per_ref, rtol=1e-10, atol=1e-12)
