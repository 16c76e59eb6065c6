"""Device assembly of the finite-difference preconditioning matrices (sb200_elliptic_jacobian_csr = FormJacobian
elliptic.C:537-590; sb200_stokes_pc_velocity_csr = StokesPCSetUp0 stokes.C:1160-1240) against the oracle's matrices about
the same state: identical sparsity pattern (bit-exact rowptr / colidx), values within 1e-12 relative (fp64; the GPU state
itself differs from the oracle's by ~1e-14 and nvcc contracts a*b+c)."""
import numpy as np
import pytest
import scipy.sparse as sps
import torch

import spectral_petsc_b200 as sp
from oracle.elliptic import MatElliptic
from oracle.stokes import StokesCtx

pytestmark = pytest.mark.gpu
TOL = 1e-12


def same_matrix(csr, ref):
    rowptr, colidx, vals = [t.cpu().numpy() for t in csr]
    ref = ref.tocsr().copy()
    ref.sort_indices()
    assert rowptr.dtype == np.int32 and colidx.dtype == np.int32
    assert np.array_equal(rowptr, ref.indptr)
    assert np.array_equal(colidx, ref.indices)
    assert np.abs(vals - ref.data).max() <= TOL * np.abs(ref.data).max()


@pytest.mark.parametrize("dim,gamma", [([8, 6], 4.0), ([7, 6, 5], 4.0), ([16, 16, 16], 0.0), ([16, 16, 16], 4.0), ([3, 3, 3], 4.0),
                                       ([12] * 5, 4.0), ([5, 4, 3, 6], 1.5), ([32, 32, 32], 4.0)], ids=str)
def test_elliptic_jacobian_csr_equals_oracle(cuda, dim, gamma):
    O = MatElliptic(dim, gamma=gamma, exponent=2.0)
    G = sp.Elliptic(dim, gamma=gamma, exponent=2.0)
    Us = 0.1 * np.random.default_rng(1).standard_normal(O.g)
    O.form_function(Us)
    G.form_function(torch.from_numpy(Us).to(cuda))
    csr = G.jacobian_csr()
    same_matrix(csr, O.form_jacobian_matrix())
    # a second state: values-only refresh into the same pattern (SAME_NONZERO_PATTERN, elliptic.C:588)
    Us2 = 0.1 * np.random.default_rng(5).standard_normal(O.g)
    O.form_function(Us2)
    G.form_function(torch.from_numpy(Us2).to(cuda))
    csr2 = G.jacobian_csr(pattern=csr[:2])
    assert csr2[0] is csr[0] and csr2[1] is csr[1]
    same_matrix(csr2, O.form_jacobian_matrix())


@pytest.mark.parametrize("dim,rheology", [([8, 6], 0), ([8, 6], 1), ([7, 6, 5], 1), ([12, 12, 12], 1), ([20, 20, 20], 0)], ids=str)
def test_stokes_velocity_pc_csr_equals_oracle(cuda, dim, rheology):
    O = StokesCtx(dim, rheology=rheology, exponent=3.0, regularization=1e-2, exact=2)
    O.create_exact_solution()
    G = sp.Stokes(dim, rheology=rheology, exponent=3.0, regularization=1e-2)
    G.set_dirichlet(torch.from_numpy(O.dirichlet.reshape(-1).copy()).to(cuda))
    G.set_force(torch.from_numpy(O.force).to(cuda))
    x = 0.3 * np.random.default_rng(2).standard_normal(O.g)
    O.function(x)
    G.function(torch.from_numpy(x).to(cuda))
    same_matrix(G.pc_velocity_csr(), O.pc_velocity_matrix())


def test_jacobian_csr_as_the_pc_input_of_config_1(cuda):
    """config 1 (elliptic 16^3 -exact 2 -ksp_rtol 1e-10): an LU of the DEVICE-assembled matrix preconditions the solve to the
    same iteration count as an LU of the oracle's matrix (13, tests/test_gpu_ksp.py)."""
    import scipy.sparse.linalg as spla

    from oracle.fgmres import fgmres

    dim = [16, 16, 16]
    O = MatElliptic(dim)
    u, u2 = O.create_exact_solution(2)
    G = sp.Elliptic(dim)
    G.set_dirichlet(torch.from_numpy(O.dirichlet).to(cuda))
    G.set_rhs(torch.from_numpy(O.b).to(cuda))
    x0 = np.zeros(O.g)
    F = G.form_function(torch.from_numpy(x0).to(cuda)).cpu().numpy()
    O.form_function(x0)
    rowptr, colidx, vals = [t.cpu().numpy() for t in G.jacobian_csr()]
    P = sps.csr_matrix((vals, colidx, rowptr), shape=(O.g, O.g))
    lu_g = spla.splu(P.tocsc())
    lu_o = spla.splu(O.form_jacobian_matrix().tocsc())
    A = lambda v: G.mat_mult(torch.from_numpy(np.ascontiguousarray(v)).to(cuda)).cpu().numpy()
    rg = fgmres(A, -F, M=lu_g.solve, rtol=1e-10)
    ro = fgmres(A, -F, M=lu_o.solve, rtol=1e-10)
    assert rg[3] == ro[3] == 2  # converged on rtol
    assert rg[1] == ro[1] == 13
    assert np.abs(rg[0] - ro[0]).max() <= 1e-9 * np.abs(ro[0]).max()


def test_large_matrix_sizes_and_slab_refusal(cuda):
    G = sp.Elliptic([128, 128, 128], gamma=4.0, exponent=2.0)
    G.form_function(torch.from_numpy(0.1 * np.random.default_rng(1).standard_normal(G.g)).to(cuda))
    rowptr, colidx, vals = G.jacobian_csr()
    n = 126
    assert rowptr.numel() == n ** 3 + 1 and vals.numel() == n ** 3 + 3 * 2 * (n - 1) * n * n
    rp = rowptr.cpu().numpy().astype(np.int64)
    assert rp[0] == 0 and rp[-1] == vals.numel() and (np.diff(rp) >= 4).all() and (np.diff(rp) <= 7).all()
    ci = colidx.cpu().numpy()
    assert ci.min() == 0 and ci.max() == n ** 3 - 1
    # row sums of the gamma-free part vanish away from the boundary; here just: finite, diagonal positive
    v = vals.cpu().numpy()
    assert np.isfinite(v).all()
    diag = v[[rp[r] + int(np.searchsorted(ci[rp[r]:rp[r + 1]], r)) for r in (0, 1000, n ** 3 // 2, n ** 3 - 1)]]
    assert (diag > 0).all()
    S = sp.Elliptic([16, 16, 16], rank=0, nranks=2)
    with pytest.raises(sp.SB200Error) as ei:
        S.jacobian_csr()
    assert ei.value.code == 56
