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


def _slices(group, epsi, epsa, T1, T2, Vpphh, Vhhhp, Vppph, vertex, v):
    """Host side of one group: the sub-tensors its engine needs (contiguous column-major copies)."""
    holes, trip, _ = group
    U = np.array(holes)
    f = np.asfortranarray
    d = dict(epsi=np.asarray(epsi)[U], T1=f(np.asarray(T1)[:, U]), T2aa=f(np.asarray(T2)[:, :, U][:, :, :, U]),
             T2al=f(np.asarray(T2)[:, :, U, :]), Vpphh=f(np.asarray(Vpphh)[:, :, U][:, :, :, U]),
             Vhhhp=f(np.asarray(Vhhhp)[U][:, U]))
    if Vppph is not None:
        d["Vppph"] = f(np.asarray(Vppph)[:, :, :, U])
    else:
        np_ = vertex.shape[1]
        sel = np.concatenate([U, np.arange(np_ - v, np_)])
        d["vertex"] = np.asarray(vertex)[:, sel][:, :, sel]
    local = {h: n for n, h in enumerate(holes)}
    lidx = {t: n for n, t in enumerate(_sorted_triples(len(holes)))}
    d["want"] = [lidx[tuple(local[h] for h in t)] for _, t in trip]
    return d


def run_out_of_core(epsi, epsa, T1, T2, Vpphh, Vhhhp, Vppph=None, vertex=None, block: int = 8,
                    device: int = 0, world: int = 1, rank: int = 0, prefetch: bool = True):
    """Returns (sum of E_t over this rank's groups, per-triple array with this rank's entries).
    With ``prefetch`` the host slicing of the next group runs in a worker thread while the GPU works
    on the current one (the C ABI calls release the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    o, v = int(len(epsi)), int(len(epsa))
    if Vppph is None and vertex is None:
        raise ValueError("Missing argument: PPPHCoulombIntegrals (or CoulombVertex)")
    per = np.zeros(o * (o + 1) * (o + 2) // 6)
    total = 0.0
    mine = deal(block_groups(o, block), world, rank)
    args = (epsi, epsa, T1, T2, Vpphh, Vhhhp, Vppph, vertex, v)
    with ThreadPoolExecutor(max_workers=1) as pool:
        nxt = pool.submit(_slices, mine[0], *args) if mine else None
        for n, (holes, trip, _) in enumerate(mine):
            d = nxt.result()
            nxt = pool.submit(_slices, mine[n + 1], *args) if (prefetch and n + 1 < len(mine)) else None
            with TriplesEngine(len(holes), v, device=device, o_all=o) as eng:
                eng.set_eigenenergies(d["epsi"], epsa)
                eng.set_singles(d["T1"])
                eng.set_doubles(d["T2aa"])
                eng.set_doubles_hole(d["T2al"])
                eng.set_pphh(d["Vpphh"])
                eng.set_hhhp(d["Vhhhp"])
                if "Vppph" in d:
                    eng.set_ppph(d["Vppph"])
                else:
                    eng.set_vertex(d["vertex"])
                res = eng.run_list(d["want"])
            if nxt is None and n + 1 < len(mine):
                nxt = pool.submit(_slices, mine[n + 1], *args)
            for (g, _), e in zip(trip, res.per_triple):
                per[g] = e
            total += res.energy
    return total, per
