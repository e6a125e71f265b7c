"""The fused kernel's step / routing tables, verified on the CPU: a tile-level NumPy
emulation of exactly the table-driven data flow the kernel executes
(tools/gen_tables.py:emulate_triple) must reproduce the oracle's per-triple energy
for every degeneracy class of the hole triple and of the particle-range orbit."""
import os

import pytest

from oracle import pt_oracle as O
from sisi4s_b200 import synthetic as S
import gen_tables as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_is_up_to_date():
    import io
    buf = io.StringIO()
    G.emit_header(buf)
    with open(os.path.join(ROOT, "sisi4s_b200", "csrc", "pt_tables.h")) as f:
        assert f.read() == buf.getvalue(), "run: python tools/gen_tables.py > sisi4s_b200/csrc/pt_tables.h"


def test_table_shapes():
    steps = [[len(G.ALL_TABLES[tc][oc]["steps"]) for oc in range(4)] for tc in range(4)]
    assert steps == [[18, 9, 9, 3], [12, 6, 6, 2], [12, 6, 6, 2], [6, 3, 3, 1]]
    assert [G.ALL_TABLES[tc][0]["distinct"] for tc in range(4)] == [[0, 1, 2, 3, 4, 5], [0, 2, 4], [0, 1, 2], [0]]
    assert G.ALL_TABLES[3][0]["coef"] == [0.0] * 6  # i=j=k contributes exactly zero
    for tc in range(4):
        for oc in range(4):
            t = G.ALL_TABLES[tc][oc]
            assert len(t["tiles"]) == [6, 3, 3, 1][oc]
            for st in t["steps"]:
                assert st["halves"][0]["en"] == 1


@pytest.mark.parametrize("kind", ["random", "vertex"])
@pytest.mark.parametrize("o,v,tile", [(3, 10, 4), (2, 7, 4), (3, 5, 8), (2, 17, 16)])
def test_emulated_fused_flow_matches_oracle(kind, o, v, tile):
    inp = S.make_inputs(o, v, seed=4, kind=kind, kappa=1.0 if kind == "vertex" else None)
    for ijk in O.sorted_triples(o):
        a = O.triple_energy(*inp.args(), ijk)
        b = G.emulate_triple(*inp.args(), ijk, tile=tile)
        assert abs(a - b) <= 1e-11 * max(1.0, abs(a)), (ijk, a, b)


def test_generic_orbit_tables_are_the_s3_group_table():
    """The symmetric epilogue of the fused kernel (generic orbit A>B>C) relies on: tile mu of the
    orbit holds the points x o mu of tile 0, and nbr[mu][nu] is the index of the composed
    permutation m -> mu(nu(m)) -- the constant S3_MUL in pt_fused.cu."""
    import re
    src = open(os.path.join(ROOT, "sisi4s_b200", "csrc", "pt_fused.cu")).read()
    m = re.search(r"S3_MUL\[6\]\[6\] = \{(.*?)\};", src, re.S)
    s3 = [[int(x) for x in row.split(",")] for row in re.findall(r"\{([0-9, ]+)\}", m.group(1))]
    P = [tuple(p) for p in G.PERM]
    mul = [[P.index(tuple(P[a][P[b][k]] for k in range(3))) for b in range(6)] for a in range(6)]
    assert s3 == mul
    for tc in range(4):
        t = G.ALL_TABLES[tc][0]
        assert len(t["tiles"]) == 6
        assert [list(r) for r in t["nbr"]] == mul
