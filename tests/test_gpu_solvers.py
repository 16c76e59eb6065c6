"""The solver orchestration over the CUDA shells + device FGMRES against the same orchestration over the CPU oracle: BASELINE
config 4 (./stokes -exact 2 ... -ksp_type fgmres -ksp_rtol 1e-10, Schur block LU, linear viscosity) and the nonlinear elliptic
Newton loop give identical outer iteration counts (+-1) and the same error against the manufactured solution (north_star)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla
import torch

import spectral_petsc_b200 as sp
from spectral_petsc_b200 import solvers
from oracle.elliptic import MatElliptic
from oracle.fgmres import fgmres
from oracle.stokes import StokesCtx

pytestmark = pytest.mark.gpu


def np_krylov(op, b, pc, rtol, maxits, restart):
    x, its, hist, reason = fgmres(op, b, M=pc, restart=restart, rtol=rtol, maxits=maxits)
    return x, its, reason


@pytest.mark.parametrize("dim,saddle", [([12, 12, 12], 0), ([16, 16, 16], 0), ([12, 12, 12], 3)], ids=lambda v: str(v))
def test_config4_stokes_block_lu_matches_oracle(cuda, dim, saddle):
    O = StokesCtx(dim, rheology=0, exact=2)
    U, _ = O.create_exact_solution()
    F0 = O.function(np.zeros(O.g))
    lu = spla.splu(O.pc_velocity_matrix().tocsc())  # the velocity PC stays outside the package (stand-in for hypre on MatVVPC)
    pco = solvers.StokesSaddlePC(O, 3, np_krylov, lu.solve, saddle_type=saddle)
    dxo, its_o, reason_o = solvers.solve_stokes_linear(O, 3, np_krylov, pco, -F0, rtol=1e-10, maxits=200)

    S = sp.Stokes(dim, rheology=0)
    S.set_dirichlet(torch.from_numpy(O.dirichlet.reshape(-1).copy()).to(cuda))
    S.set_force(torch.from_numpy(O.force).to(cuda))
    F = S.function(torch.zeros(S.g, dtype=torch.float64, device=cuda))
    gk = solvers.make_gpu_krylov()
    vpc = lambda r: torch.from_numpy(lu.solve(r.cpu().numpy())).to(cuda)
    pcg = solvers.StokesSaddlePC(S, 3, gk, vpc, saddle_type=saddle)
    dx, its, reason = solvers.solve_stokes_linear(S, 3, gk, pcg, -1.0 * F, rtol=1e-10, maxits=200)
    assert reason == reason_o == 2
    assert abs(its - its_o) <= 1
    dx = dx.cpu().numpy()
    v, p = solvers.split(dx, 3)
    vo, po = solvers.split(dxo, 3)
    ve, _ = solvers.split(U, 3)
    err, err_o = np.abs(v - ve).max(), np.abs(vo - ve).max()
    assert abs(err - err_o) < 1e-8 + 1e-3 * err_o  # the same "Norm of error"
    assert np.abs(v - vo).max() < 1e-7 * max(np.abs(vo).max(), 1.0)
    assert abs(pcg.inner_its["velocity"] - pco.inner_its["velocity"]) <= 2 + pco.inner_its["velocity"] // 20


def test_newton_elliptic_nonlinear_matches_oracle(cuda):
    dim = [16, 16]
    O = MatElliptic(dim, gamma=4.0, exponent=2.0)
    u, _ = O.create_exact_solution(0, cos_scale=1.0)

    def jac_o(rhs):
        lu = spla.splu(O.form_jacobian_matrix().tocsc())
        x, its, _ = np_krylov(O.mat_mult, rhs, lu.solve, 1e-12, 200, 30)
        return x, its

    xo, its_o, kits_o, hist_o = solvers.newton(O.form_function, jac_o, np.zeros(O.g), rtol=1e-12)

    G = sp.Elliptic(dim, gamma=4.0, exponent=2.0)
    G.set_dirichlet(torch.from_numpy(O.dirichlet).to(cuda))
    G.set_rhs(torch.from_numpy(O.b).to(cuda))
    H = MatElliptic(dim, gamma=4.0, exponent=2.0)  # host-side FormJacobian input: the state the GPU residual cached
    gk = solvers.make_gpu_krylov()

    def jac_g(rhs):
        H.eta, H.deta = G.get_state(0).cpu().numpy(), G.get_state(1).cpu().numpy()
        H.gradu = [G.get_state(2 + k).cpu().numpy() for k in range(H.d)]
        lu = spla.splu(H.form_jacobian_matrix().tocsc())
        x, its, _ = gk(G.mat_mult, rhs, lambda r: torch.from_numpy(lu.solve(r.cpu().numpy())).to(cuda), 1e-12, 200, 30)
        return x, its

    x, its, kits, hist = solvers.newton(lambda v: G.form_function(v).clone(), jac_g, torch.zeros(G.g, dtype=torch.float64, device=cuda), rtol=1e-12)
    assert abs(its - its_o) <= 1  # SNES iteration count
    assert all(abs(a - b) <= 1 for a, b in zip(kits, kits_o))  # KSP iteration counts per Newton step
    assert np.abs(x.cpu().numpy() - u).max() < 1e-9 and abs(np.abs(x.cpu().numpy() - u).max() - np.abs(xo - u).max()) < 1e-10


def test_config5_power_law_continuation_matches_oracle(cuda):
    """./stokes -exact 2 -cont 4 -rheology 1 -eps 1e-4 -exponent 3 (README:55, BASELINE config 5) at a small extent: the five SNES
    solves of the continuation take the same numbers of Newton and Krylov iterations on the CUDA shells as on the oracle."""
    dim = [8, 8, 8]
    kw = dict(rheology=1, hardness=1.0, exponent=3.0, regularization=1e-4, gamma0=1.0)
    O = StokesCtx(dim, exact=2, **kw)
    O.create_exact_solution()

    def pc_o():
        return solvers.StokesSaddlePC(O, 3, np_krylov, spla.splu(O.pc_velocity_matrix().tocsc()).solve, saddle_type=0)

    xo, log_o = solvers.solve_stokes_continuation(O.function, O.set_rheology, pc_o, O, 3, np_krylov, np.zeros(O.g), 3.0, 1e-4, cont=4,
                                                  ksp_rtol=1e-6, snes_rtol=1e-8, ksp_maxits=300)

    S = sp.Stokes(dim, **kw)
    S.set_dirichlet(torch.from_numpy(O.dirichlet.reshape(-1).copy()).to(cuda))
    S.set_force(torch.from_numpy(O.force).to(cuda))
    H = StokesCtx(dim, exact=2, **kw)  # host-side StokesPCSetUp0 input: the viscosity the GPU residual cached
    gk = solvers.make_gpu_krylov()

    def pc_g():
        H.eta = S.get_state(0).cpu().numpy()
        lu = spla.splu(H.pc_velocity_matrix().tocsc())
        return solvers.StokesSaddlePC(S, 3, gk, lambda r: torch.from_numpy(lu.solve(r.cpu().numpy())).to(cuda), saddle_type=0)

    x, log = solvers.solve_stokes_continuation(lambda v: S.function(v).clone(), lambda e, r: S.set_rheology(1, 1.0, e, r, 1.0), pc_g, S, 3, gk,
                                               torch.zeros(S.g, dtype=torch.float64, device=cuda), 3.0, 1e-4, cont=4,
                                               ksp_rtol=1e-6, snes_rtol=1e-8, ksp_maxits=300)
    assert len(log) == len(log_o) == 5
    for a, b in zip(log, log_o):
        assert abs(a["snes_its"] - b["snes_its"]) <= 1, (a, b)
        for ka, kb in zip(a["ksp_its"], b["ksp_its"]):
            assert abs(ka - kb) <= 1 + kb // 10, (a, b)
    xv, _ = solvers.split(x.cpu().numpy(), 3)
    xov, _ = solvers.split(xo, 3)
    assert np.abs(xv - xov).max() < 1e-6 * max(np.abs(xov).max(), 1.0)
    mn, mx = S.eta_minmax()
    assert mn == pytest.approx(O.min_eta, rel=1e-6) and mx == pytest.approx(O.max_eta, rel=1e-6)
