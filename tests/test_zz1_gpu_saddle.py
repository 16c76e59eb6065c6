"""(Runs late in the GPU suite: written after this round's GPU budget was spent; CPU-verified over the test double.)
The device-resident saddle-point preconditioners StokesPCApply0..3 (sb200_saddle_*, host/saddle.cpp + csrc/vecops.cu) against
the same composition written with torch operations (spectral_petsc_b200.solvers.StokesSaddlePC over the CUDA shells, which
tests/test_gpu_solvers.py pins to the oracle flow), and the vector helpers against torch."""
import numpy as np
import pytest
import torch

import spectral_petsc_b200 as sp
from spectral_petsc_b200 import solvers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nodes,d", [(1, 2), (1000, 3), (126 ** 3, 3), (77, 2), (300001, 1)])
def test_vector_helpers_equal_torch(cuda, nodes, d):
    g = torch.Generator(device="cpu").manual_seed(nodes + d)
    x = torch.randn(nodes * (d + 1), dtype=torch.float64, generator=g).to(cuda)
    v, p = sp.vec_split(x, d)
    X = x.reshape(nodes, d + 1)
    assert torch.equal(v, X[:, :d].reshape(-1)) and torch.equal(p, X[:, d].reshape(-1))
    assert torch.equal(sp.vec_merge(v, p, d), x)
    y = torch.randn(nodes * d, dtype=torch.float64, generator=g).to(cuda)
    for a, b in ((1.0, -1.0), (2.5, 0.0), (0.0, -1.0), (-0.75, 3.0)):
        ref = a * v + b * y if a != 0.0 else b * y
        out = sp.vec_axpby(a, v if a != 0.0 else None, b, y.clone())
        assert torch.allclose(out, ref, rtol=1e-14, atol=1e-13)  # one fused multiply-add may round differently
    out = sp.vec_axpby(2.0, v, 0.0, torch.full_like(v, float("nan")))  # b == 0 must not read y
    assert torch.equal(out, 2.0 * v)
    diag = torch.rand(nodes, dtype=torch.float64, generator=g).to(cuda) + 0.5
    assert torch.equal(sp.vec_pointwise_divide(p, diag), p / diag)
    q = p.clone() + 3.0
    sp.vec_remove_mean(q)
    assert abs(float(q.mean())) < 1e-12 and torch.allclose(q, p + 3.0 - (p + 3.0).mean(), rtol=0, atol=1e-11)
    z = x.clone()
    sp.vec_remove_mean(z, stride=d + 1, offset=d)  # the pressure slots of the global vector (stokes.C:1013-1023)
    Z = z.reshape(nodes, d + 1)
    assert torch.equal(Z[:, :d], X[:, :d]) and torch.allclose(Z[:, d], X[:, d] - X[:, d].mean(), rtol=0, atol=1e-11)
    z2 = x.clone()
    sp.vec_remove_mean(z2, stride=d + 1, offset=d)
    assert torch.equal(z, z2)  # deterministic reduction order


def _state(cuda, dim, rheology):
    U, U2, dirichlet = sp.stokes_exact_solution(dim, 2)
    S = sp.Stokes(dim, rheology=rheology, exponent=2.0 if rheology else 1.0, regularization=0.5 if rheology else 1.0)
    S.set_dirichlet(torch.from_numpy(dirichlet.reshape(-1).copy()).to(cuda))
    S.set_force(torch.from_numpy(U2).to(cuda))
    S.function(torch.from_numpy(U).to(cuda))  # eta, deta, strain of the manufactured solution
    return S


@pytest.mark.parametrize("saddle", [0, 1, 2, 3])
@pytest.mark.parametrize("dim,rheology,preonly", [([8, 8, 8], 0, True), ([10, 9, 8], 1, True), ([8, 8, 8], 1, False), ([12, 10], 1, True)], ids=str)
def test_device_saddle_equals_torch_composition(cuda, saddle, dim, rheology, preonly):
    d = len(dim)
    S = _state(cuda, dim, rheology)
    rowptr, colidx, vals = [t.cpu().numpy() for t in S.pc_velocity_csr()]
    import scipy.sparse as sps

    dinv = torch.from_numpy(1.0 / sps.csr_matrix((vals, colidx, rowptr), shape=(S.gv, S.gv)).diagonal()).to(cuda)
    vpc = lambda r: dinv * r  # Jacobi on MatVVPC stands for PETSc's PC (stays on the device: no host round trip in this test)
    ref = solvers.StokesSaddlePC(S, d, solvers.make_gpu_krylov(), vpc, saddle_type=saddle, vel_max_it=4, schur_max_it=3, svel_preonly=preonly, svel_max_it=4)
    dev = sp.StokesSaddle(S, saddle, velocity_pc=vpc, vel_max_it=4, schur_max_it=3, svel_preonly=preonly, svel_max_it=4)  # (-svel_ksp_max_it 4)
    x = torch.from_numpy(np.random.default_rng(saddle).standard_normal(S.g)).to(cuda)
    y_ref = ref.apply(x)
    y = dev.apply(x)
    scale = float(y_ref.abs().max())
    assert float((y - y_ref).abs().max()) <= 1e-10 * scale
    assert all(abs(dev.inner_its[k] - ref.inner_its[k]) <= 1 for k in ("velocity", "schur")) and dev.inner_its["velocity"] > 0
    y2 = dev.apply(x, remove_constant_pressure=True)
    assert float((y2 - solvers.remove_constant_pressure(y_ref, d)).abs().max()) <= 1e-10 * scale
    with pytest.raises(sp.SB200Error):
        dev.apply(x, x)
    dev.destroy()
    S.destroy()


def test_block_lu_with_exact_inner_solves_inverts_the_operator(cuda):
    """stokes.C:1712-1713: "If applied exactly, this is a direct method." """
    S = _state(cuda, [7, 7, 7], 1)
    dev = sp.StokesSaddle(S, 0, velocity_pc=None, vel_rtol=1e-12, vel_max_it=2000, schur_rtol=1e-12, schur_max_it=500, svel_preonly=False, svel_rtol=1e-12, svel_max_it=2000)
    x = torch.from_numpy(np.random.default_rng(5).standard_normal(S.g)).to(cuda)
    sp.vec_remove_mean(x, stride=4, offset=3)
    y = dev.apply(S.mat_mult(x), remove_constant_pressure=True)
    assert float((y - x).abs().max()) < 1e-8 * float(x.abs().max())
    dev.destroy()
    S.destroy()


def test_outer_fgmres_with_native_saddle_pc_matches_python_pc(cuda):
    """BASELINE config 4 shape at 12^3: the outer FGMRES with the PC applied natively (no Python in the iteration except the
    velocity-PC stand-in) takes the same iterations to the same solution as with the torch composition."""
    import scipy.sparse as sps
    import scipy.sparse.linalg as spla

    dim, d = [12, 12, 12], 3
    U, U2, dirichlet = sp.stokes_exact_solution(dim, 2)
    S = sp.Stokes(dim, rheology=0)
    S.set_dirichlet(torch.from_numpy(dirichlet.reshape(-1).copy()).to(cuda))
    S.set_force(torch.from_numpy(U2).to(cuda))
    F = S.function(torch.zeros(S.g, dtype=torch.float64, device=cuda))
    rowptr, colidx, vals = [t.cpu().numpy() for t in S.pc_velocity_csr()]
    lu = spla.splu(sps.csr_matrix((vals, colidx, rowptr), shape=(S.gv, S.gv)).tocsc())
    vpc = lambda r: torch.from_numpy(lu.solve(r.cpu().numpy())).to(cuda)
    gk = solvers.make_gpu_krylov()
    ref = solvers.StokesSaddlePC(S, d, gk, vpc, saddle_type=0)
    dx_ref, its_ref, reason_ref = solvers.solve_stokes_linear(S, d, gk, ref, -1.0 * F, rtol=1e-10, maxits=200)
    dev = sp.StokesSaddle(S, 0, velocity_pc=vpc, vel_max_it=4, schur_max_it=3, svel_preonly=True)
    K = sp.KSP(S.g)
    K.set_operators(S, pc=dev)
    K.set_tolerances(rtol=1e-10, maxits=200)
    dx = K.solve(-1.0 * F)
    assert K.result["reason"] == reason_ref == 2 and abs(K.result["its"] - its_ref) <= 1
    assert float((dx - dx_ref).abs().max()) < 1e-7 * float(dx_ref.abs().max())
    ve = torch.from_numpy(U).to(cuda).reshape(-1, 4)[:, :3]
    err, err_ref = float((dx.reshape(-1, 4)[:, :3] - ve).abs().max()), float((dx_ref.reshape(-1, 4)[:, :3] - ve).abs().max())
    assert abs(err - err_ref) < 1e-8 + 1e-3 * err_ref  # the same "Norm of error" against the manufactured solution
    K.destroy()
    dev.destroy()
    S.destroy()


def test_device_saddle_hits_the_oracle_golden_vectors(cuda):
    """tests/golden/saddle_7x6x5.npz (tests/golden/make_golden.py: StokesPCApply0..3 composed over the ORACLE shells and the oracle's
    FGMRES): the device composition over the CUDA shells reproduces those vectors and the inner iteration counts."""
    import os

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "saddle_7x6x5.npz"))
    dim = [int(v) for v in z["dim"]]
    S = _state(cuda, dim, 1)
    import scipy.sparse as sps

    rowptr, colidx, vals = [t.cpu().numpy() for t in S.pc_velocity_csr()]
    dinv = torch.from_numpy(1.0 / sps.csr_matrix((vals, colidx, rowptr), shape=(S.gv, S.gv)).diagonal()).to(cuda)
    x = torch.from_numpy(z["x"]).to(cuda)
    for t in range(4):
        dev = sp.StokesSaddle(S, t, velocity_pc=lambda r: dinv * r, vel_max_it=4, schur_max_it=3, svel_preonly=True)
        y = dev.apply(x).cpu().numpy()
        assert np.abs(y - z["y%d" % t]).max() <= 1e-9 * np.abs(z["y%d" % t]).max()
        assert [dev.inner_its["velocity"], dev.inner_its["schur"]] == [int(v) for v in z["its%d" % t]]
        dev.destroy()
    S.destroy()


@pytest.mark.parametrize("dim", [[8, 6], [9, 7, 6], [32, 32, 32], [6, 5, 4, 3]], ids=str)
def test_csr_diagonal_of_the_device_assembled_matrices(cuda, dim):
    """sb200_csr_diagonal (MatGetDiagonal for PCJacobi, no download of the matrix) on FormJacobian's P and, in 2-D / 3-D, on MatVVPC."""
    def by_torch(rowptr, colidx, vals):
        counts = (rowptr[1:] - rowptr[:-1]).long()
        rows = torch.repeat_interleave(torch.arange(rowptr.numel() - 1, device=rowptr.device), counts)
        return vals[colidx.long() == rows]

    E = sp.Elliptic(dim, gamma=4.0, exponent=2.0)
    E.form_function(torch.from_numpy(0.1 * np.random.default_rng(1).standard_normal(E.g)).to(cuda))
    csr = E.jacobian_csr()
    d = sp.csr_diagonal(*csr)
    assert d.numel() == E.g and torch.equal(d, by_torch(*csr)) and float(d.abs().min()) > 0.0
    E.destroy()
    if len(dim) in (2, 3):
        S = _state(cuda, dim, 1)
        csr = S.pc_velocity_csr()
        d = sp.csr_diagonal(*csr)
        assert d.numel() == S.gv and torch.equal(d, by_torch(*csr)) and float(d.abs().min()) > 0.0
        S.destroy()
