"""The saddle-point preconditioners (StokesPCApply0..3, stokes.C:1714-1817) and an outer FGMRES solve on a SLAB-partitioned Stokes
context (BASELINE config 5 is "stokes ... slab-partitioned over 2/4/8"): ranks emulated on one device, one host thread and one stream
per rank, arenas mapped by pointer.  The inner Krylov solvers' dot products and the constant-pressure mean cross the ranks through
peer memory; the result must be the single-GPU result (same iteration counts; vectors to rounding)."""
import threading

import numpy as np
import pytest
import torch

import spectral_petsc_b200 as sp
from spectral_petsc_b200 import dist as spd
from conftest import no_gc_during_collective

pytestmark = pytest.mark.gpu


def _state(cuda, dim, rheology, rank=0, nranks=1):
    U, U2, dirichlet = sp.stokes_exact_solution(dim, 2)
    d = len(dim)
    S = sp.Stokes(dim, rheology=rheology, exponent=2.0 if rheology else 1.0, regularization=0.5 if rheology else 1.0, rank=rank, nranks=nranks)
    return S, U, U2, dirichlet.reshape(-1)


def _run_ranks(nr, fn):
    out, err = [None] * nr, [None] * nr

    def run(r):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                out[r] = fn(r)
                torch.cuda.current_stream().synchronize()
        except Exception as e:  # pragma: no cover
            err[r] = e

    with no_gc_during_collective():
        torch.cuda.synchronize()
        th = [threading.Thread(target=run, args=(r,)) for r in range(nr)]
        for t in th:
            t.start()
        for t in th:
            t.join(timeout=300)
    assert all(e is None for e in err), err
    assert all(o is not None for o in out)
    return out


@pytest.mark.parametrize("dim,nr,rheology,saddle,preonly", [([12, 12, 12], 2, 1, 0, True), ([16, 16, 16], 4, 0, 0, False), ([12, 12, 12], 2, 1, 1, True),
                                                             ([12, 10], 2, 1, 2, True), ([12, 12, 12], 3, 0, 3, True)], ids=str)
def test_slab_saddle_apply_equals_single_gpu(cuda, dim, nr, rheology, saddle, preonly):
    d = len(dim)
    S1, U, U2, dirichlet = _state(cuda, dim, rheology)
    S1.set_dirichlet(torch.from_numpy(dirichlet.copy()).to(cuda))
    S1.set_force(torch.from_numpy(U2).to(cuda))
    S1.function(torch.from_numpy(U).to(cuda))
    kw = dict(vel_max_it=4, schur_max_it=3, svel_preonly=preonly, svel_max_it=4)
    ref = sp.StokesSaddle(S1, saddle, velocity_pc=None, **kw)
    x = np.random.default_rng(saddle).standard_normal(S1.g)
    y_ref = ref.apply(torch.from_numpy(x).to(cuda), remove_constant_pressure=True).cpu().numpy()
    its_ref = dict(ref.inner_its)

    ctx = [_state(cuda, dim, rheology, r, nr)[0] for r in range(nr)]
    spd.attach_in_process(ctx)
    dl = spd.split_dirichlet(dirichlet, dim, nr, ncomp=d)
    fl = spd.split_global(U2, dim, nr, ncomp=d + 1)
    ul = spd.split_global(U, dim, nr, ncomp=d + 1)
    xl = spd.split_global(x, dim, nr, ncomp=d + 1)
    for r, c in enumerate(ctx):
        c.set_dirichlet(torch.from_numpy(dl[r].copy()).to(cuda))
        c.set_force(torch.from_numpy(fl[r].copy()).to(cuda))
    torch.cuda.synchronize()
    _run_ranks(nr, lambda r: ctx[r].function(torch.from_numpy(ul[r].copy()).to(cuda)))  # eta / deta / strain of the state (collective)
    pcs = [sp.StokesSaddle(c, saddle, velocity_pc=None, **kw) for c in ctx]
    spd.attach_in_process(pcs)
    torch.cuda.synchronize()
    ys = _run_ranks(nr, lambda r: pcs[r].apply(torch.from_numpy(xl[r].copy()).to(cuda), remove_constant_pressure=True).cpu().numpy())
    y = np.concatenate(ys)
    assert np.abs(y - y_ref).max() <= 1e-9 * np.abs(y_ref).max()
    for p in pcs:
        assert p.inner_its == its_ref  # every rank takes the single-GPU iteration counts
    assert all(c.slab_timeouts() == 0 for c in ctx)
    for p in pcs:
        p.destroy()
    ref.destroy()


def test_slab_outer_fgmres_with_saddle_pc(cuda):
    """Config 4 / 5 shape: outer FGMRES(30) with the block-LU saddle PC (Jacobi of the finite-difference velocity matrix as the velocity
    PC), everything slab-partitioned over 2 ranks: the same iteration count and solution as on one GPU."""
    dim, d, nr = [12, 12, 12], 3, 2
    kw = dict(vel_max_it=4, schur_max_it=3, svel_preonly=True)

    def solve(S, pc, K, rhs):
        K.set_operators(S, pc=pc)
        K.set_tolerances(rtol=1e-8, maxits=300)
        x = K.solve(rhs)
        return x.cpu().numpy(), K.result

    S1, U, U2, dirichlet = _state(cuda, dim, 0)
    S1.set_dirichlet(torch.from_numpy(dirichlet.copy()).to(cuda))
    S1.set_force(torch.from_numpy(U2).to(cuda))
    F = S1.function(torch.zeros(S1.g, dtype=torch.float64, device=cuda))
    rhs = (-1.0 * F).cpu().numpy()
    diag = sp.csr_diagonal(*S1.pc_velocity_csr())  # MatGetDiagonal(MatVVPC): the Jacobi stand-in for PETSc's PC
    pc1 = sp.StokesSaddle(S1, 0, velocity_pc=lambda r: r / diag, **kw)
    x_ref, r_ref = solve(S1, pc1, sp.KSP(S1.g), torch.from_numpy(rhs).to(cuda))
    assert r_ref["reason"] == 2, r_ref

    ctx = [_state(cuda, dim, 0, r, nr)[0] for r in range(nr)]
    spd.attach_in_process(ctx)
    dl = spd.split_dirichlet(dirichlet, dim, nr, ncomp=d)
    fl = spd.split_global(U2, dim, nr, ncomp=d + 1)
    rl = spd.split_global(rhs, dim, nr, ncomp=d + 1)
    dg = [torch.from_numpy(a.copy()).to(cuda) for a in spd.split_global(diag.cpu().numpy(), dim, nr, ncomp=d)]
    for r, c in enumerate(ctx):
        c.set_dirichlet(torch.from_numpy(dl[r].copy()).to(cuda))
        c.set_force(torch.from_numpy(fl[r].copy()).to(cuda))
    torch.cuda.synchronize()
    _run_ranks(nr, lambda r: ctx[r].function(torch.zeros(ctx[r].g, dtype=torch.float64, device=cuda)))
    pcs = [sp.StokesSaddle(c, 0, velocity_pc=(lambda v, r=r: v / dg[r]), **kw) for r, c in enumerate(ctx)]
    spd.attach_in_process(pcs)
    ksps = [sp.KSP(c.g, rank=r, nranks=nr) for r, c in enumerate(ctx)]
    spd.attach_in_process(ksps)
    torch.cuda.synchronize()
    res = _run_ranks(nr, lambda r: solve(ctx[r], pcs[r], ksps[r], torch.from_numpy(rl[r].copy()).to(cuda)))
    x = np.concatenate([a for a, _ in res])
    assert all(rr["reason"] == 2 and abs(rr["its"] - r_ref["its"]) <= 1 for _, rr in res), (r_ref, [rr for _, rr in res])
    assert res[0][1]["its"] == res[1][1]["its"]
    assert np.abs(x - x_ref).max() <= 1e-3 * np.abs(x_ref).max()  # (two solves to rtol 1e-8 of an ill-conditioned system)
    assert all(c.slab_timeouts() == 0 for c in ctx)
