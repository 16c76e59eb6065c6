"""Host-side logic of the slab partition on CPU: world_size-2 (and 4) gloo process groups, no GPU.

Covers the partition arithmetic (Python mirror vs the C ABI's sb200_slab_geometry), the scatter / gather of
global Vecs and Dirichlet vectors against the oracle's index maps, and the handle exchange used to map the
peers' arenas (the 64-byte CUDA IPC handles are replaced by rank-stamped dummies)."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spectral_petsc_b200 import dist as spd  # noqa: E402
import spectral_petsc_b200 as sp  # noqa: E402
from oracle.elliptic import MatElliptic  # noqa: E402


def c_geometry(dim, rank, nranks):
    L = sp.lib()
    arr = (ctypes.c_int * len(dim))(*dim)
    i0, nloc = ctypes.c_int(), ctypes.c_int()
    goff, g, m, nd = (ctypes.c_longlong() for _ in range(4))
    rc = L.sb200_slab_geometry(len(dim), arr, rank, nranks, ctypes.byref(i0), ctypes.byref(nloc), ctypes.byref(goff), ctypes.byref(g),
                               ctypes.byref(m), ctypes.byref(nd))
    return rc, (i0.value, nloc.value, goff.value, g.value), m.value, nd.value


@pytest.mark.parametrize("dim,nranks", [([8, 6], 2), ([16, 16, 16], 4), ([128, 128, 128], 8), ([12, 5, 4, 6], 3), ([12] * 5, 2), ([16, 16, 16], 8)])
def test_partition_arithmetic_matches_c_abi_and_oracle(dim, nranks):
    O = MatElliptic(dim) if np.prod(dim) <= 300000 else None
    gsum = ndsum = msum = 0
    for r in range(nranks):
        rc, geo, m, nd = c_geometry(dim, r, nranks)
        assert rc == 0
        assert geo == spd.slab_range(dim, r, nranks)
        assert nd == spd.dirichlet_range(dim, r, nranks)[1]
        i0, nloc, goff, g = geo
        plane = int(np.prod(dim[1:]))
        if O is not None:
            # the local Vec range is exactly the interior nodes whose flat index lies in this rank's planes
            mine = np.flatnonzero((O.ixG >= i0 * plane) & (O.ixG < (i0 + nloc) * plane))
            assert mine.size == g and (g == 0 or (mine[0] == goff and mine[-1] == goff + g - 1))
            dmine = np.flatnonzero((O.ixD >= i0 * plane) & (O.ixD < (i0 + nloc) * plane))
            doff, ndl = spd.dirichlet_range(dim, r, nranks)
            assert dmine.size == ndl and (ndl == 0 or dmine[0] == doff)
        gsum += g
        ndsum += nd
        msum += m
    assert msum == int(np.prod(dim)) and gsum == int(np.prod([p - 2 for p in dim])) and gsum + ndsum == msum


def test_partition_errors():
    assert c_geometry([10, 8, 8], 0, 4)[0] == 83   # not divisible
    assert c_geometry([16, 16], 2, 2)[0] == 83     # rank out of range
    assert c_geometry([16, 16], 0, 9)[0] == 83     # more than 8 ranks
    with pytest.raises(ValueError):
        spd.slab_range([10, 8], 0, 4)


def _worker(rank, world, port, dim, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = int(np.prod([p - 2 for p in dim]))
        U = np.random.default_rng(0).standard_normal(g)
        local = torch.from_numpy(spd.split_global(U, dim, world)[rank].copy())
        back = spd.gather_global(local).numpy()
        ok_vec = np.array_equal(back, U)
        V3 = np.random.default_rng(1).standard_normal(4 * g)  # Stokes AoS [v, p]: 4 values per interior node
        loc3 = torch.from_numpy(spd.split_global(V3, dim, world, ncomp=4)[rank].copy())
        ok_aos = np.array_equal(spd.gather_global(loc3).numpy(), V3)
        handle = bytes([rank]) * 64
        hs = spd.exchange_handles(handle)
        ok_h = [h == bytes([r]) * 64 for r, h in enumerate(hs)]
        q.put((rank, ok_vec, ok_aos, all(ok_h), len(hs)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dim", [(2, [16, 16, 16]), (4, [8, 6, 5]), (2, [12] * 4)])
def test_gloo_scatter_gather_and_handle_exchange(world, dim):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + len(dim)
    procs = [ctx.Process(target=_worker, args=(r, world, port, dim, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == list(range(world))
    for _, ok_vec, ok_aos, ok_h, n in res:
        assert ok_vec and ok_aos and ok_h and n == world
