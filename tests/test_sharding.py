"""N>1 host logic on CPU: world_size-2 `gloo` ranks shard the sorted triples with
pt_partition and all-reduce the scalar energy (sisi4s_b200/sharding.py).  The per-range
evaluator here is the C oracle (the CUDA engine needs a GPU; the same TripleShards object
drives it in bench.py and in the 2-GPU test below)."""
import os
import socket

import numpy as np
import pytest

from sisi4s_b200 import synthetic as S
from sisi4s_b200.sharding import TripleShards, partition

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "triples_golden.npz"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, o, v, nbatch, out):
    import torch.distributed as dist
    from oracle import c_oracle as CO
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        inp = S.make_inputs(o, v, seed=2026, kind="vertex")
        sh = TripleShards(o)
        assert (sh.world, sh.rank) == (world, rank)
        seen = []

        def evaluate(lo, hi):
            seen.append((lo, hi))
            return float(CO.triples_list(*inp.args(), np.arange(lo, hi), nthreads=1).sum())

        e = sh.run(evaluate, nbatch=nbatch)
        ntr = sh.sum(float(sum(hi - lo for lo, hi in seen)))
        tmax = sh.max(float(rank))
        if rank == 0:
            out.put((e, ntr, tmax))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nbatch", [1, 3])
def test_two_gloo_ranks_reproduce_golden_energy(nbatch):
    import torch.multiprocessing as mp
    o, v, world = 5, 19, 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, o, v, nbatch, out)) for r in range(world)]
    for p in procs:
        p.start()
    e, ntr, tmax = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert abs(e - float(GOLDEN["o5_v19_vertex_s2026_total"])) <= 1e-12
    assert ntr == o * (o + 1) * (o + 2) // 6   # every sorted triple exactly once
    assert tmax == world - 1


@pytest.mark.parametrize("o,world,nbatch", [(5, 2, 1), (20, 8, 1), (40, 8, 8), (64, 8, 1), (3, 4, 2)])
def test_rank_ranges_tile_the_triple_space(o, world, nbatch):
    ntr = o * (o + 1) * (o + 2) // 6
    w = np.array([[6, 3, 3, 1][(i == j) + 2 * (j == k)] for i in range(o) for j in range(i, o) for k in range(j, o)])
    cur = 0
    loads = []
    for b in range(nbatch):
        for r in range(world):
            lo, hi = TripleShards(o, world, r).my_range(nbatch, b)
            assert lo == cur and hi >= lo
            cur = hi
            loads.append(int(w[lo:hi].sum()))
    assert cur == ntr
    assert (lo, hi) == partition(o, nbatch * world, nbatch * world - 1)
    if ntr >= 50 * nbatch * world:   # equal weight up to one triple per cut
        assert max(loads) - min(loads) <= 12


def _nccl_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from sisi4s_b200.triples import TriplesEngine
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        inp = S.make_inputs(4, 33, seed=17, kind="vertex")
        sh = TripleShards(4)
        with TriplesEngine(4, 33, device=rank) as eng:
            eng.set_inputs(*inp.args())
            e = sh.run(lambda lo, hi: eng.run(lo, hi).energy, nbatch=2, device=torch.device("cuda", rank))
        if rank == 0:
            out.put(e)
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpus_reproduce_golden_energy():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    e = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert abs(e - float(GOLDEN["o4_v33_vertex_s17_total"])) <= 1e-9
