"""Slab partition of the Stokes shells (BASELINE config 5 layout) on ONE device: ranks are separate contexts on
separate streams with arenas mapped by pointer (see tests/test_gpu_slab.py).  Every rank's local result must equal
the matching slice of the oracle's single-domain result; bar 1e-12 max-norm relative."""
import numpy as np
import pytest
import torch

import spectral_petsc_b200 as sp
from spectral_petsc_b200 import dist as spd
from oracle.stokes import StokesCtx
from conftest import rel_max, no_gc_during_collective

pytestmark = pytest.mark.gpu
TOL = 1e-12


def setup(dim, nranks, rheology, cuda, exponent=3.0, eps=1e-2):
    O = StokesCtx(dim, rheology=rheology, hardness=1.0, exponent=exponent, regularization=eps, gamma0=1.0, exact=2)
    O.create_exact_solution()
    d = len(dim)
    ctx = [sp.Stokes(dim, rheology=rheology, hardness=1.0, exponent=exponent, regularization=eps, gamma0=1.0, rank=r, nranks=nranks)
           for r in range(nranks)]
    spd.attach_in_process(ctx)
    streams = [torch.cuda.Stream(device=cuda) for _ in range(nranks)]
    assert sum(c.g for c in ctx) == O.g and sum(c.gp for c in ctx) == O.gp and sum(c.dv for c in ctx) == O.dv
    dl = spd.split_dirichlet(O.dirichlet.reshape(-1), dim, nranks, ncomp=d)
    fl = spd.split_global(O.force, dim, nranks, ncomp=d + 1)
    torch.cuda.synchronize()
    for r, c in enumerate(ctx):
        with torch.cuda.stream(streams[r]):
            c.set_dirichlet(torch.from_numpy(dl[r].copy()).to(cuda))
            c.set_force(torch.from_numpy(fl[r].copy()).to(cuda))
    torch.cuda.synchronize()
    return O, ctx, streams


def each(ctx, streams, fn):
    out = []
    with no_gc_during_collective():
        torch.cuda.synchronize()
        for r, c in enumerate(ctx):
            with torch.cuda.stream(streams[r]):
                out.append(fn(r, c))
        torch.cuda.synchronize()
    return np.concatenate([o.cpu().numpy() for o in out])


# [16..], [32..] with equal extents take the pencil (transpose) path for axis 0 + batched even-odd launches; the others
# the operand-pull path with the generic kernels
CASES = [([8, 6], 2, 0), ([8, 6], 2, 1), ([12, 7, 6], 2, 1), ([12, 7, 6], 3, 0), ([16, 16, 16], 4, 1), ([16, 10, 12], 8, 1), ([20, 20, 20], 2, 1),
         ([16, 16], 2, 1), ([32, 32, 32], 2, 1), ([32, 32, 32], 8, 0), ([16, 16, 16], 8, 1)]


@pytest.mark.parametrize("dim,nranks,rheology", CASES, ids=lambda v: str(v))
def test_stokes_slab_matches_oracle(cuda, dim, nranks, rheology):
    O, ctx, streams = setup(dim, nranks, rheology, cuda)
    d = len(dim)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    xs = 0.3 * np.random.default_rng(1).standard_normal(O.g)
    parts = spd.split_global(xs, dim, nranks, ncomp=d + 1)
    F = each(ctx, streams, lambda r, c: c.function(dev(parts[r])))
    assert rel_max(F, O.function(xs)) < TOL
    mn = min(c.eta_minmax()[0] for c in ctx)
    mx = max(c.eta_minmax()[1] for c in ctx)
    assert mn == pytest.approx(O.min_eta, rel=1e-13) and mx == pytest.approx(O.max_eta, rel=1e-13)
    x = np.random.default_rng(0).standard_normal(O.g)
    v, p = O.split(x)
    xp = spd.split_global(x, dim, nranks, ncomp=d + 1)
    vp = spd.split_global(v, dim, nranks, ncomp=d)
    pp = spd.split_global(p, dim, nranks, ncomp=1)
    assert rel_max(each(ctx, streams, lambda r, c: c.mat_mult(dev(xp[r]))), O.mat_mult(x)) < TOL
    assert rel_max(each(ctx, streams, lambda r, c: c.mat_mult_vv(dev(vp[r]))), O.mat_mult_vv(v)) < TOL
    assert rel_max(each(ctx, streams, lambda r, c: c.mat_mult_pv(dev(vp[r]))), O.mat_mult_pv(v)) < TOL
    assert rel_max(each(ctx, streams, lambda r, c: c.mat_mult_vp(dev(pp[r]))), O.mat_mult_vp(p)) < TOL
    assert rel_max(each(ctx, streams, lambda r, c: c.get_diagonal_schur()), O.get_diagonal_schur()) < 1e-13
    assert all(c.slab_timeouts() == 0 for c in ctx)


def test_stokes_slab_null_space(cuda):
    # stokes.C:206-212: the constant pressure is in the null space of the partitioned operator too
    dim, nr = [16, 16, 16], 4
    O, ctx, streams = setup(dim, nr, 0, cuda)
    ns = O.merge(np.zeros(O.gv), np.ones(O.gp))
    parts = spd.split_global(ns, dim, nr, ncomp=4)
    y = each(ctx, streams, lambda r, c: c.mat_mult(torch.from_numpy(parts[r].copy()).to(cuda)))
    assert np.abs(y).max() < 1e-12
