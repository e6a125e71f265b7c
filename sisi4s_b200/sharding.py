"""Multi-GPU sharding of the (T) step: one process per GPU, the sorted hole triples
(i<=j<=k, reference enumeration order, src/algorithms/CcsdPerturbativeTriples.cxx:156-158)
cut into contiguous chunks of equal weight (pt_partition, weight = number of distinct hole
permutations = W blocks to build: 6/3/3/1), inputs replicated, and ONE collective: the
all-reduce of the scalar energy.  In the reference the same sum is the CTF::Scalar
accumulation of :214 (an MPI all-reduce per permutation of every triple).

The module only holds host logic (ranges + the collective); it runs unchanged over
`gloo` on CPU ranks (tests/test_sharding.py) and `nccl` on GPUs (bench.py).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable

from . import _lib


def partition(o: int, nparts: int, part: int) -> tuple[int, int]:
    """[begin, end) of part `part` of `nparts` (host-only C ABI call, no GPU needed)."""
    b, e = C.c_int64(), C.c_int64()
    _lib.check(_lib.load().pt_partition(int(o), int(nparts), int(part), C.byref(b), C.byref(e)))
    return int(b.value), int(e.value)


class TripleShards:
    """This rank's view of the triple space within a torch.distributed group.

    With ``nbatch`` > 1 the list is first cut into ``nbatch`` weight-balanced batches
    (bench.py's "steps") and every batch is split over the ranks, so rank r owns part
    ``batch * world + r`` of ``nbatch * world``.
    """

    def __init__(self, o: int, world: int | None = None, rank: int | None = None, group=None):
        self.o = int(o)
        self.group = group
        if world is None or rank is None:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                world, rank = dist.get_world_size(group), dist.get_rank(group)
            else:
                world, rank = 1, 0
        self.world, self.rank = int(world), int(rank)

    def my_range(self, nbatch: int = 1, batch: int = 0) -> tuple[int, int]:
        return partition(self.o, nbatch * self.world, (batch % nbatch) * self.world + self.rank)

    # -- the collectives: one double each
    def _reduce(self, x: float, op: str, device=None) -> float:
        if self.world == 1:
            return float(x)
        import torch
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op), group=self.group)
        return float(t.item())

    def sum(self, x: float, device=None) -> float:
        return self._reduce(x, "SUM", device)

    def max(self, x: float, device=None) -> float:
        return self._reduce(x, "MAX", device)

    def run(self, evaluate: Callable[[int, int], float], nbatch: int = 1, device=None) -> float:
        """E(T) = all-reduce of sum over this rank's ranges of evaluate(begin, end)."""
        local = 0.0
        for b in range(nbatch):
            lo, hi = self.my_range(nbatch, b)
            if hi > lo:
                local += float(evaluate(lo, hi))
        return self.sum(local, device)
