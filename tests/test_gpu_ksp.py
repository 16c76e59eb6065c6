"""Device-resident FGMRES (C ABI sb200_ksp_*) against the FGMRES oracle: identical iteration counts (+-1), the same
residual history and the same error against the -exact manufactured solution (BASELINE north_star)."""
import threading

import numpy as np
import pytest
import scipy.sparse.linalg as spla
import torch

import spectral_petsc_b200 as sp
from spectral_petsc_b200 import dist as spd
from oracle.elliptic import MatElliptic
from oracle.fgmres import fgmres
from conftest import rel_max, no_gc_during_collective

pytestmark = pytest.mark.gpu


def test_dense_system_matches_oracle(cuda):
    rng = np.random.default_rng(0)
    n = 300
    A = rng.standard_normal((n, n)) + 2.0 * n ** 0.5 * np.eye(n)
    b = rng.standard_normal(n)
    Ad = torch.from_numpy(A).to(cuda)
    for restart in (30, 7):
        xo, its_o, hist_o, reason_o = fgmres(lambda v: A @ v, b, restart=restart, rtol=1e-10)
        K = sp.KSP(n, restart=restart)
        K.set_operators(lambda v: Ad @ v)
        K.set_tolerances(rtol=1e-10)
        x = K.solve(torch.from_numpy(b).to(cuda)).cpu().numpy()
        r = K.result
        assert r["reason"] == reason_o == 2
        assert abs(r["its"] - its_o) <= 1
        h = K.history
        k = min(len(h), len(hist_o)) - 2
        assert np.allclose(h[:k], hist_o[:k], rtol=1e-6)
        assert np.linalg.norm(A @ x - b) <= 2e-10 * np.linalg.norm(b)
        assert rel_max(x, xo) < 1e-7


def test_config1_elliptic16_exact2_ksp_rtol_1e10(cuda):
    """./elliptic -dim 16,16,16 -exact 2 -ksp_rtol 1e-10: one Newton step of the linear problem from x = 0."""
    dim = [16, 16, 16]
    O = MatElliptic(dim, gamma=0.0)
    u, _ = O.create_exact_solution(2)
    F0 = O.form_function(np.zeros(O.g))
    lu = spla.splu(O.form_jacobian_matrix().tocsc())  # stand-in for PETSc's PC on the FD matrix (out of scope)
    dxo, its_o, hist_o, _ = fgmres(O.mat_mult, -F0, M=lu.solve, rtol=1e-10)

    G = sp.Elliptic(dim, gamma=0.0)
    G.set_dirichlet(torch.from_numpy(O.dirichlet).to(cuda))
    G.set_rhs(torch.from_numpy(O.b).to(cuda))
    F = G.form_function(torch.zeros(G.g, dtype=torch.float64, device=cuda))
    assert rel_max(F.cpu().numpy(), F0) < 1e-12
    K = sp.KSP(G.g)
    K.set_operators(G, pc=lambda r: torch.from_numpy(lu.solve(r.cpu().numpy())).to(cuda))
    K.set_tolerances(rtol=1e-10)
    dx = K.solve(-F).cpu().numpy()
    r = K.result
    assert r["reason"] == 2 and abs(r["its"] - its_o) <= 1
    assert np.allclose(K.history[:its_o - 1], hist_o[:its_o - 1], rtol=1e-5)
    err, err_o = np.abs(dx - u).max(), np.abs(dxo - u).max()
    assert err < 1e-9 and abs(err - err_o) < 1e-10  # "Norm of error" (elliptic.C:214-226)
    t = K.times_ms
    assert t["operator"] > 0 and t["pc"] > 0 and t["ksp_vector_work"] > 0


def test_native_operator_no_pc_restarts(cuda):
    dim = [8, 8, 8]
    O = MatElliptic(dim, gamma=4.0, exponent=2.0)
    O.create_exact_solution(2)
    Us = 0.1 * np.random.default_rng(1).standard_normal(O.g)
    O.form_function(Us)
    b = np.random.default_rng(0).standard_normal(O.g)
    xo, its_o, hist_o, reason_o = fgmres(O.mat_mult, b, restart=30, rtol=1e-8, maxits=400)
    G = sp.Elliptic(dim, gamma=4.0, exponent=2.0)
    G.set_dirichlet(torch.from_numpy(O.dirichlet).to(cuda))
    G.set_rhs(torch.from_numpy(O.b).to(cuda))
    G.form_function(torch.from_numpy(Us).to(cuda))
    K = sp.KSP(G.g)
    K.set_operators(G)
    K.set_tolerances(rtol=1e-8, maxits=400)
    x = K.solve(torch.from_numpy(b).to(cuda)).cpu().numpy()
    r = K.result
    assert r["reason"] == reason_o
    assert abs(r["its"] - its_o) <= max(1, its_o // 50)  # long restarted runs may drift by an iteration or two
    if reason_o == 2:
        assert np.linalg.norm(O.mat_mult(x) - b) <= 1e-7 * np.linalg.norm(b)


def test_slab_ksp_two_ranks_in_process(cuda):
    """Vectors and operator slab-partitioned over 2 ranks (emulated on one device, one host thread per rank):
    the dot products go through the peer-memory all-reduce; counts and solution equal the single-domain solve."""
    import os

    os.environ["SB200_MAX_CTAS"] = "6"
    try:
        dim, nr = [16, 16, 16], 2
        O = MatElliptic(dim, gamma=0.0)
        O.create_exact_solution(2)
        b = np.random.default_rng(0).standard_normal(O.g)
        xo, its_o, hist_o, reason_o = fgmres(O.mat_mult, b, rtol=1e-6, maxits=90)
        ctx = [sp.Elliptic(dim, gamma=0.0, rank=r, nranks=nr) for r in range(nr)]
        spd.attach_in_process(ctx)
        ksp = [sp.KSP(c.g, rank=r, nranks=nr) for r, c in enumerate(ctx)]
        spd.attach_in_process(ksp)
        parts = spd.split_global(b, dim, nr)
        out, res = [None] * nr, [None] * nr

        def run(r):
            with torch.cuda.stream(torch.cuda.Stream(device=cuda)):
                ksp[r].set_operators(ctx[r])
                ksp[r].set_tolerances(rtol=1e-6, maxits=90)
                out[r] = ksp[r].solve(torch.from_numpy(parts[r].copy()).to(cuda)).cpu().numpy()
                res[r] = ksp[r].result

        with no_gc_during_collective():
            torch.cuda.synchronize()
            th = [threading.Thread(target=run, args=(r,)) for r in range(nr)]
            for t in th:
                t.start()
            for t in th:
                t.join(timeout=120)
        assert all(o is not None for o in out)
        assert res[0]["its"] == res[1]["its"] and res[0]["rnorm"] == res[1]["rnorm"]  # same bits on every rank
        assert res[0]["reason"] == reason_o and abs(res[0]["its"] - its_o) <= 1
        x = np.concatenate(out)
        assert rel_max(x, xo) < 1e-5
        assert all(c.slab_timeouts() == 0 for c in ctx)
    finally:
        del os.environ["SB200_MAX_CTAS"]


@pytest.mark.parametrize("restart,maxits", [(30, 10000), (7, 10000), (30, 17)])
def test_lookahead_gives_the_same_solve(cuda, restart, maxits):
    """sb200_ksp_set_lookahead(1): step k+1 is enqueued before the norm of step k is read.  Same iterates bit for bit, same
    iteration count, history and reason - whether the solve ends by convergence (a speculative step is discarded), at a restart
    boundary or at the iteration cap; the operator is applied at most once more."""
    rng = np.random.default_rng(0)
    n = 400
    A = rng.standard_normal((n, n)) + 2.0 * n ** 0.5 * np.eye(n)
    Ad = torch.from_numpy(A).to(cuda)
    b = torch.from_numpy(rng.standard_normal(n)).to(cuda)
    out = []
    for la in (0, 1):
        calls = [0]

        def op(v):
            calls[0] += 1
            return Ad @ v

        K = sp.KSP(n, restart=restart)
        K.set_operators(op, pc=lambda r: 0.5 * r)
        K.set_tolerances(rtol=1e-10, maxits=maxits)
        K.set_lookahead(la)
        x = K.solve(b).clone()
        out.append((x, dict(K.result), list(K.history), calls[0]))
        K.destroy()
    (x0, r0, h0, c0), (x1, r1, h1, c1) = out
    assert torch.equal(x0, x1)
    assert r0["its"] == r1["its"] and r0["reason"] == r1["reason"] and h0 == h1
    assert c0 <= c1 <= c0 + 1
    with pytest.raises(sp.SB200Error):
        K2 = sp.KSP(8)
        K2.set_lookahead(2)
