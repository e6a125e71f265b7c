#!/usr/bin/env python3
"""Generate the step / routing tables of the fused (T) kernel and emulate its
data flow on the CPU.

The fused kernel (sisi4s_b200/csrc/pt_fused.cu) processes one *work item* =
(sorted hole triple i<=j<=k, orbit {A>=B>=C} of 16-wide particle ranges).  For
that item it runs a short list of *stacked GEMM steps*; each step multiplies up
to two 16-row T2 panels against ONE shared 256-pair tile of a PPPH slab, and
adds the two resulting 16^3 W tiles into the on-chip X tiles of the orbit with
an index permutation.  Afterwards the epilogue combines the X tiles of the
orbit (six index permutations), the singles term and the eigenvalue
denominator into the item's energy.  All the index algebra that decides *which*
panels, *which* X tile and *which* permutation lives in the tables generated
here, so it can be verified on the CPU (tests/test_tables.py) against the
oracle before any GPU is involved.

Derivation (reference CcsdPerturbativeTriples.cxx:159-216):
  h = (i,j,k) sorted, pi_p = Permutation<3>(p), h_p = h o pi_p.
  distinct p: first occurrence of each h_p (reference :165-177).
  Xd[x]  = sum_{p distinct} W_{h_p}[x o pi_p]              x = (a,b,c)
  X[x]   = sum_{all p} W_{h_p}[x o pi_p] = sum_{g in Stab(h)} Xd[x o g]
  Sd[x]  = sum_{p distinct} S_{h_p}[x o pi_p]
  E_ijk  = sum_x X[x]/D[x] * sum_s sf(s) (Xd+Sd)[x o sigma_s]
         = sum_x (Xd+Sd)[x] * ( sum_nu c_nu Xd[x o nu] ) / D[x]
  with c_nu = sum_{g in Stab(h)} sf(nu o g^-1)   (substitute x -> x o sigma;
  D is symmetric; sf is a class function).

Usage:  python tools/gen_tables.py > sisi4s_b200/csrc/pt_tables.h
"""
from __future__ import annotations

import itertools
import sys

import numpy as np

# Permutation<3>(p).images, reference src/math/Permutation.hpp:52-62; pinned
# against the reference header by oracle/ref_perm_dump.cxx (tests/test_oracle.py)
PERM = [(0, 1, 2), (1, 0, 2), (1, 2, 0), (0, 2, 1), (2, 0, 1), (2, 1, 0)]
SPIN_AND_FERMI = (+2.0, -4.0, 0.0, +8.0)  # CcsdPerturbativeTriples.cxx:143
TILE = 16


def after(f, tau):
    return tuple(f[tau[m]] for m in range(3))


def inverse(pi):
    inv = [0, 0, 0]
    for m in range(3):
        inv[pi[m]] = m
    return tuple(inv)


def sf(sigma):
    return SPIN_AND_FERMI[sum(1 for m in range(3) if sigma[m] == m)]


# class representatives: slots with equal values share a canonical slot
TRIPLE_CLASS_CANON = [(0, 1, 2), (0, 0, 2), (0, 1, 1), (0, 0, 0)]   # i<j<k, i=j<k, i<j=k, i=j=k
ORBIT_CLASS_CANON = [(0, 1, 2), (0, 0, 2), (0, 1, 1), (0, 0, 0)]    # A>B>C, A=B>C, A>B=C, A=B=C


def triple_class(i, j, k):
    return (1 if i == j else 0) + (2 if j == k else 0)


def build_class_tables(tc: int, oc: int):
    hc = TRIPLE_CLASS_CANON[tc]
    rc = ORBIT_CLASS_CANON[oc]
    # distinct hole permutations, first occurrence (reference :165-177)
    distinct = []
    for p in range(6):
        if all(after(hc, PERM[q]) != after(hc, PERM[p]) for q in range(p)):
            distinct.append(p)
    stab = [g for g in PERM if after(hc, g) == hc]
    coef = []
    for nu in PERM:
        coef.append(sum(sf(after(nu, inverse(g))) for g in stab))
    # X tiles: distinct arrangements of the orbit's ranges
    tiles = []          # slot tuples (first representative)
    tile_key = {}
    for rho in PERM:
        key = after(rc, rho)
        if key not in tile_key:
            tile_key[key] = len(tiles)
            tiles.append(rho)
    nbr = [[tile_key[after(rc, after(alpha, nu))] for nu in PERM] for alpha in tiles]
    # group distinct p by the VALUE of the slab hole z = h_p[2], pair them up
    groups = {}
    for p in distinct:
        z = after(hc, PERM[p])[2]
        groups.setdefault(z, []).append(p)
    pairs = []
    for z in sorted(groups, reverse=True):
        ps = groups[z]
        for n in range(0, len(ps), 2):
            pairs.append(tuple(ps[n:n + 2]))
    steps = []
    for pair in pairs:
        for rho in tiles:        # W tile arrangement (a-range; b-range, c-range)
            halves = []
            for p in pair:
                pi = PERM[p]
                q = inverse(pi)
                # X-tile arrangement: x_n range = rg[rho[q[n]]]
                tau = tile_key[after(rc, after(rho, q))]
                halves.append(dict(en=1, p=p, tx=pi[0], ty=pi[1], tau=tau, q=q))
            zs = PERM[pair[0]][2]
            while len(halves) < 2:
                halves.append(dict(en=0, p=0, tx=0, ty=0, tau=0, q=(0, 1, 2)))
            steps.append(dict(zs=zs, rho=rho, halves=halves))
    pmask = sum(1 << p for p in distinct)
    return dict(tc=tc, oc=oc, distinct=distinct, pmask=pmask, coef=coef,
                tiles=tiles, nbr=nbr, steps=steps)


ALL_TABLES = [[build_class_tables(tc, oc) for oc in range(4)] for tc in range(4)]


# ---------------------------------------------------------------------------
# CPU emulation of the fused data flow (tile-level; dense numpy per tile)
# ---------------------------------------------------------------------------
def emulate_triple(epsi, epsa, T1, T2, Vpphh, Vhhhp, Vppph, ijk, tile=TILE):
    """Energy contribution of one sorted triple, following exactly the table
    driven flow of the fused kernel (zero padding to multiples of `tile`)."""
    o, v = epsi.size, epsa.size
    nr = (v + tile - 1) // tile
    vp = nr * tile
    h = tuple(ijk)
    tc = triple_class(*h)

    def pad(x, axes):
        padw = [(0, 0)] * x.ndim
        for ax in axes:
            padw[ax] = (0, vp - v)
        return np.pad(x, padw)

    T2p = pad(T2, (0, 1))
    Vp = pad(Vppph, (0, 1, 2))
    Up = pad(Vhhhp, (3,))
    Pp = pad(Vpphh, (0, 1))
    T1p = pad(T1, (0,))
    eap = np.concatenate([epsa, np.zeros(vp - v)])

    def rng(r):
        return slice(r * tile, (r + 1) * tile)

    e_total = 0.0
    for A in range(nr):
        for B in range(A + 1):
            for C in range(B + 1):
                rg = (A, B, C)
                oc = (1 if A == B else 0) + (2 if B == C else 0)
                tab = ALL_TABLES[tc][oc]
                X = np.zeros((len(tab["tiles"]), tile, tile, tile))
                for st in tab["steps"]:
                    r0, r1, r2 = (rg[s] for s in st["rho"])
                    z = h[st["zs"]]
                    Vt = Vp[rng(r1), rng(r2), :, z]            # [b,c,d]
                    for hf in st["halves"]:
                        if not hf["en"]:
                            continue
                        x_, y_ = h[hf["tx"]], h[hf["ty"]]
                        # W[a,b,c] = sum_d T2[a,d,x,y] V[b,c,d,z] - sum_l T2[a,b,x,l] U[y,z,l,c]
                        W = np.einsum("ad,bcd->abc", T2p[rng(r0), :v, x_, y_], Vt[:, :, :v])
                        W -= np.einsum("abl,lc->abc", T2p[rng(r0), rng(r1), x_, :], Up[y_, z, :, rng(r2)])
                        # X_tau[x] += W[w], x_n = w_{q[n]}  ->  X axis n is W axis q[n]
                        X[hf["tau"]] += W.transpose(hf["q"])
                # epilogue
                e3 = epsi[h[0]] + epsi[h[1]] + epsi[h[2]]
                pm = tab["pmask"]
                for t, alpha in enumerate(tab["tiles"]):
                    ra, rb, rc_ = (rg[s] for s in alpha)
                    ga = np.arange(ra * tile, (ra + 1) * tile)
                    gb = np.arange(rb * tile, (rb + 1) * tile)
                    gc = np.arange(rc_ * tile, (rc_ + 1) * tile)
                    valid = ((ga < v)[:, None, None] & (gb < v)[None, :, None] & (gc < v)[None, None, :])
                    D = e3 - eap[ga][:, None, None] - eap[gb][None, :, None] - eap[gc][None, None, :]
                    Z = np.zeros((tile,) * 3)
                    for n, nu in enumerate(PERM):
                        # Xd[x o nu]: value at coords (x_nu0,x_nu1,x_nu2) of tile nbr; as array over x:
                        # arr[x0,x1,x2] = Xn[x_nu0, x_nu1, x_nu2] -> Xn axis m is x axis nu[m]
                        Xn = X[tab["nbr"][t][n]]
                        Z += tab["coef"][n] * np.transpose(Xn, inverse(nu))
                    i_, j_, k_ = h
                    Qa = ((pm >> 0 & 1) * Pp[rb * tile:(rb + 1) * tile, rc_ * tile:(rc_ + 1) * tile, j_, k_]
                          + (pm >> 3 & 1) * Pp[rc_ * tile:(rc_ + 1) * tile, rb * tile:(rb + 1) * tile, k_, j_].T)
                    Qb = ((pm >> 1 & 1) * Pp[ra * tile:(ra + 1) * tile, rc_ * tile:(rc_ + 1) * tile, i_, k_]
                          + (pm >> 2 & 1) * Pp[rc_ * tile:(rc_ + 1) * tile, ra * tile:(ra + 1) * tile, k_, i_].T)
                    Qc = ((pm >> 4 & 1) * Pp[ra * tile:(ra + 1) * tile, rb * tile:(rb + 1) * tile, i_, j_]
                          + (pm >> 5 & 1) * Pp[rb * tile:(rb + 1) * tile, ra * tile:(ra + 1) * tile, j_, i_].T)
                    ta, tb, tcv = T1p[ga, i_], T1p[gb, j_], T1p[gc, k_]
                    Sd = 0.5 * (ta[:, None, None] * Qa[None, :, :]
                                + tb[None, :, None] * Qb[:, None, :]
                                + tcv[None, None, :] * Qc[:, :, None])
                    contrib = np.where(valid, (X[t] + Sd) * Z / np.where(valid, D, 1.0), 0.0)
                    e_total += float(contrib.sum())
    return e_total


# ---------------------------------------------------------------------------
# header emission
# ---------------------------------------------------------------------------
def emit_header(out=sys.stdout):
    w = out.write
    w("// GENERATED by tools/gen_tables.py -- do not edit.\n")
    w("// Step / routing tables of the fused (T) kernel; verified on the CPU against the\n")
    w("// oracle by tests/test_tables.py (emulate_triple).\n")
    w("#pragma once\n#include <cstdint>\n\n")
    w("struct PtHalf { int8_t en, p, tx, ty, tau, q0, q1, q2; };\n")
    w("struct PtStep { int8_t zs, r0, r1, r2; PtHalf h[2]; };\n")
    w("struct PtClassTable {\n  int32_t nsteps, ntiles, pmask, pad_;\n  double coef[6];\n"
      "  int8_t tile_slots[6][4];\n  int8_t nbr[6][8];\n  PtStep steps[18];\n};\n\n")
    w("#define PT_TABLES_INIT { \\\n")
    for tc in range(4):
        w("  { \\\n")
        for oc in range(4):
            t = ALL_TABLES[tc][oc]
            w("    { %d, %d, %d, 0, {" % (len(t["steps"]), len(t["tiles"]), t["pmask"]))
            w(", ".join("%.1f" % c for c in t["coef"]))
            w("}, \\\n      {")
            slots = [list(a) + [0] for a in t["tiles"]] + [[0, 0, 0, 0]] * (6 - len(t["tiles"]))
            w(", ".join("{%d,%d,%d,%d}" % tuple(s) for s in slots))
            w("}, \\\n      {")
            nb = [list(r) + [0, 0] for r in t["nbr"]] + [[0] * 8] * (6 - len(t["nbr"]))
            w(", ".join("{" + ",".join(str(x) for x in r) + "}" for r in nb))
            w("}, \\\n      {")
            ss = []
            for st in t["steps"]:
                hs = []
                for hf in st["halves"]:
                    hs.append("{%d,%d,%d,%d,%d,%d,%d,%d}" % (hf["en"], hf["p"], hf["tx"], hf["ty"],
                                                            hf["tau"], *hf["q"]))
                ss.append("{%d,%d,%d,%d,{%s}}" % (st["zs"], *st["rho"], ",".join(hs)))
            ss += ["{0,0,0,0,{{0,0,0,0,0,0,1,2},{0,0,0,0,0,0,1,2}}}"] * (18 - len(ss))
            w(", \\\n       ".join(ss))
            w("} }%s \\\n" % ("," if oc < 3 else ""))
        w("  }%s \\\n" % ("," if tc < 3 else ""))
    w("}\n")


if __name__ == "__main__":
    emit_header()
