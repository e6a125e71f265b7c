"""Synthetic inputs of the large benchmark shapes, generated ON THE DEVICE (bench / test plumbing).

`synthetic.make_inputs(kind="vertex")` builds its integrals with NumPy on the host, which takes minutes at
o=64, v=512 and beyond.  `generate_inputs` produces the same family (same vertex, same formulas:
CoulombIntegralsFromVertex.cxx:402-403, 416-417, 430-431; T2 = Vpphh / D2; T1 ~ rms(T2)) with torch FP64
einsums over the small auxiliary index on the GPU, one hole slab at a time, written straight into
page-locked host memory (`HostBuffers`).  Input generation only -- nothing of the (T) computation happens
here, and the values differ from the host generator's in the last bits (summation order), so callers
must use ONE of the two generators for a given check.
"""
from __future__ import annotations

import os
import sys

import numpy as np

NF_SYNTH = 24  # auxiliary index of the synthetic benchmark vertex (setup cost only; not on the timed path)


# host buffers of the large tensors: page-locked, private per rank or -- for the shapes whose tensors would
# not fit N times into the host -- ONE copy per node in /dev/shm shared by the ranks
class HostBuffers:
    def __init__(self, shared: bool, local: int, barrier, tag: str):
        self.shared, self.local, self.barrier, self.tag = shared, local, barrier, tag
        self.owners, self.paths, self.registered, self.allocated = [], [], [], []

    def array(self, name: str, shape):
        """Fortran-ordered float64 array in host memory: page-locked when private; a shared /dev/shm
        mapping is left pageable here (the library page-locks it read-only itself: option pin_host)."""
        import torch
        n = int(np.prod(shape))
        if not self.shared:
            # cudaHostAlloc of the exact size (torch's pinned allocator rounds up to 2^k bytes; memory that is
            # only cudaHostRegister'ed was measured slower to DMA from with 8 ranks on one host)
            a = self._host_alloc(n)
            if a is None:
                t = torch.empty(n, dtype=torch.float64, pin_memory=True)
                self.owners.append(t)
                a = t.numpy()
            return a.reshape(shape, order="F")
        path = f"/dev/shm/sisi4s_bench_{self.tag}_{name}"
        if self.local == 0:
            with open(path, "wb") as f:
                f.truncate(n * 8)
        self.barrier()
        mm = np.memmap(path, dtype=np.float64, mode="r+", shape=(n,))
        self.owners.append(mm)
        self.paths.append(path)
        return mm.reshape(shape, order="F")

    def _host_alloc(self, n):
        import ctypes
        try:
            rt = ctypes.CDLL("libcudart.so.12")
            ptr = ctypes.c_void_p()
            rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
            if rt.cudaHostAlloc(ctypes.byref(ptr), n * 8, 0) != 0 or not ptr.value:
                rt.cudaGetLastError()
                return None
            self.allocated.append((rt, ptr))
            buf = (ctypes.c_double * n).from_address(ptr.value)
            return np.frombuffer(buf, dtype=np.float64)
        except (OSError, AttributeError):
            return None

    def close(self):
        import torch
        for rt, ptr in self.allocated:
            rt.cudaFreeHost.argtypes = [__import__("ctypes").c_void_p]
            rt.cudaFreeHost(ptr)
        self.allocated.clear()
        for p in self.registered:
            torch.cuda.cudart().cudaHostUnregister(p)
        self.owners.clear()
        self.barrier()
        if self.shared and self.local == 0:
            for p in self.paths:
                try:
                    os.remove(p)
                except OSError:
                    pass


def generate_inputs(wl, dev, host: HostBuffers, rank: int, world: int, seed: int = 2026):
    """Synthetic inputs of sisi4s_b200.synthetic.make_inputs(kind="vertex"), with the large tensors
    built on the GPU (torch FP64 einsum over the NF = 24 auxiliary index: setup plumbing, untimed) and
    written straight into page-locked host memory, one hole slab at a time.  Every rank evaluates every
    slab (so scalars derived from them are bitwise the same on all ranks) but, when the host copy is
    shared, stores only its own share of the slabs."""
    import torch
    from . import synthetic as S
    o, v = wl["o"], wl["v"]
    gamma = S.make_vertex(o, v, seed, NF_SYNTH)
    epsi, epsa = S.eigenenergies(o, v)
    np_ = o + v
    a0 = np_ - v
    parts = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (gamma.real, gamma.imag)]   # [F,p,q]
    ei, ea = torch.from_numpy(epsi).to(dev), torch.from_numpy(epsa).to(dev)
    want_pphh = wl["mode"] != "hole_block"          # config 5 rebuilds PPHH from the vertex per group
    T2 = host.array("T2", (v, v, o, o))
    Vpphh = host.array("Vpphh", (v, v, o, o)) if want_pphh else None
    Vhhhp = host.array("Vhhhp", (o, o, o, v))
    Vppph = host.array("Vppph", (v, v, v, o)) if wl["ppph_host"] else None
    mine = (lambda j: True) if not host.shared else (lambda j: j % world == rank)

    def contract(spec, ia, ib):
        return sum(torch.einsum(spec, ia(g), ib(g)) for g in parts)

    sq = torch.zeros((), dtype=torch.float64, device=dev)
    ccsd = torch.zeros((), dtype=torch.float64, device=dev)
    for j in range(o):
        # Vabij[a,b,i,j] = G[G,a,i] G[G,b,j] (CoulombIntegralsFromVertex.cxx:402-403), as [i,b,a] = column-major [a,b,i]
        vj = contract("Fai,Fb->iba", lambda g: g[:, a0:, :o], lambda g: g[:, a0:, j])
        d2 = ei[:, None, None] + ei[j] - ea[None, :, None] - ea[None, None, :]
        tj = vj / d2
        sq += (tj * tj).sum()
        ccsd += ((2.0 * vj - vj.transpose(1, 2)) * tj).sum()
        if mine(j):
            torch.from_numpy(T2[:, :, :, j].T).copy_(tj)
            if want_pphh:
                torch.from_numpy(Vpphh[:, :, :, j].T).copy_(vj)
    # Vijka[i,j,k,a] = G[G,i,k] G[G,a,j] (:416-417), as [a,k,j,i]
    if mine(0):
        torch.from_numpy(Vhhhp.T).copy_(contract("Fik,Faj->akji", lambda g: g[:, :o, :o], lambda g: g[:, a0:, :o]))
    if Vppph is not None:
        for k in range(o):
            if mine(k):
                # Vabci[a,b,c,i] = G[G,a,c] G[G,b,i] (:430-431), as [c,b,a]
                torch.from_numpy(Vppph[:, :, :, k].T).copy_(
                    contract("Fac,Fb->cba", lambda g: g[:, a0:, a0:], lambda g: g[:, a0:, k]))
    torch.cuda.synchronize(dev)
    rms = float(torch.sqrt(sq / (float(v) * v * o * o)).item())
    T1 = np.asfortranarray(S.normal(seed, 3, (v, o)) * rms)
    host.barrier()
    del parts
    torch.cuda.empty_cache()
    return S.TriplesInputs(o, v, epsi, epsa, T1, T2, Vpphh, Vhhhp, Vppph, gamma, float(ccsd.item()))


