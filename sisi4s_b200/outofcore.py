"""Out-of-core driver of the (T) step for shapes whose hole-indexed tensors exceed one GPU
(BASELINE.json configs[4]: o=100, v=800 -- T2, its second packing and PPHH are 51 GB each, PPPH
410 GB).

The full tensors stay in HOST memory (numpy arrays or memory maps of TENS files).  The sorted hole
triples are grouped by hole-block triples (I<=J<=K), blocks of ``block`` consecutive holes; for each
group a sub-engine is created over the <= 3*block ACTIVE holes of those blocks (``pt_create_ex``:
the hole contraction sum_l still runs over all o holes), fed with the slices it needs

    eps_i[U], T1[:,U], T2[:,:,U,U], T2[:,:,U,:], Vpphh[:,:,U,U], Vhhhp[U,U,:,:], Vppph[:,:,:,U]
    (or the vertex rows of U + the particles, from which the engine builds its PPPH slabs)

and runs exactly the group's triples (``pt_run_list``).  E_t of a triple is the same number as in
the all-resident run, so the groups simply add up; with several ranks the groups are dealt by
weight and one scalar is all-reduced (sisi4s_b200.sharding.TripleShards.sum).  In the reference
the same memory problem is what SlicedCtfTensor / per-triple vertex products address
(src/algorithms/CcsdPerturbativeTriples.cxx:32-79,87-92).
"""
from __future__ import annotations

import itertools

import numpy as np

from .triples import TriplesEngine

_W = (6, 3, 3, 1)


def _sorted_triples(o):
    return [(i, j, k) for i in range(o) for j in range(i, o) for k in range(j, o)]


def block_groups(o: int, block: int):
    """[(holes U ascending, [(global triple index, (i,j,k)) ...], weight)] per hole-block triple."""
    index = {t: n for n, t in enumerate(_sorted_triples(o))}
    nb = (o + block - 1) // block
    rng = lambda b: range(b * block, min(o, (b + 1) * block))
    groups = []
    for I, J, K in itertools.combinations_with_replacement(range(nb), 3):
        trip = [(i, j, k) for i in rng(I) for j in rng(J) for k in rng(K) if i <= j <= k]
        if not trip:
            continue
        holes = sorted(set(rng(I)) | set(rng(J)) | set(rng(K)))
        w = sum(_W[(i == j) + 2 * (j == k)] for i, j, k in trip)
        groups.append((holes, [(index[t], t) for t in trip], w))
    return groups


def deal(groups, world: int, rank: int):
    """Longest-processing-time dealing of the groups to ranks; returns this rank's groups."""
    load = [0] * world
    mine = []
    for g in sorted(groups, key=lambda g: -g[2]):
        r = load.index(min(load))
        load[r] += g[2]
        if r == rank:
            mine.append(g)
    return mine


def run_out_of_core(epsi, epsa, T1, T2, Vpphh, Vhhhp, Vppph=None, vertex=None, block: int = 8,
                    device: int = 0, world: int = 1, rank: int = 0):
    """Returns (sum of E_t over this rank's groups, per-triple array with this rank's entries)."""
    o, v = int(len(epsi)), int(len(epsa))
    if Vppph is None and vertex is None:
        raise ValueError("Missing argument: PPPHCoulombIntegrals (or CoulombVertex)")
    per = np.zeros(o * (o + 1) * (o + 2) // 6)
    total = 0.0
    for holes, trip, _ in deal(block_groups(o, block), world, rank):
        U = np.array(holes)
        local = {h: n for n, h in enumerate(holes)}
        lidx = {t: n for n, t in enumerate(_sorted_triples(len(holes)))}
        with TriplesEngine(len(holes), v, device=device, o_all=o) as eng:
            eng.set_eigenenergies(np.asarray(epsi)[U], epsa)
            eng.set_singles(np.asarray(T1)[:, U])
            eng.set_doubles(np.asarray(T2)[:, :, U][:, :, :, U])
            eng.set_doubles_hole(np.asarray(T2)[:, :, U, :])
            eng.set_pphh(np.asarray(Vpphh)[:, :, U][:, :, :, U])
            eng.set_hhhp(np.asarray(Vhhhp)[U][:, U])
            if Vppph is not None:
                eng.set_ppph(np.asarray(Vppph)[:, :, :, U])
            else:
                np_ = vertex.shape[1]
                sel = np.concatenate([U, np.arange(np_ - v, np_)])
                eng.set_vertex(np.asarray(vertex)[:, sel][:, :, sel])
            want = [lidx[tuple(local[h] for h in t)] for _, t in trip]
            res = eng.run_list(want)
        for (g, _), e in zip(trip, res.per_triple):
            per[g] = e
        total += res.energy
    return total, per
