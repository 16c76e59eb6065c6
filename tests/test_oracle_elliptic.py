"""Pins the oracle's elliptic.C restatement against the reference's manufactured solutions
(elliptic.C:594-677) and DOF counts (elliptic.C:424)."""
import numpy as np
import pytest

from oracle.elliptic import MatElliptic


@pytest.mark.parametrize("dim,g,nd,tol", [([16, 16, 16], 2744, 1352, 5e-12), ([12] * 5, 100000, 148832, 2e-12), ([20] * 3, 5832, 2168, 2e-11)])
def test_K3_exact2_residual(dim, g, nd, tol):
    A = MatElliptic(dim)
    assert (A.g, A.nd) == (g, nd)  # K8
    u, u2 = A.create_exact_solution(2)
    r = A.form_function(u)
    assert np.abs(r).max() < tol  # "Norm of exact residual" at roundoff


def test_K4_exact1_homogeneous():
    A = MatElliptic([10, 9])
    u, u2 = A.create_exact_solution(1)
    assert np.all(A.dirichlet == 0.0)
    assert np.abs(A.form_function(u)).max() < 1e-11


def test_K4_exact0_nonlinear_converges_spectrally():
    # tests.sh: -exact 0 -cos_scale 3 / 2.8 -gamma 4: residual of the exact solution decays spectrally with n
    for cs in (3.0, 2.8):
        errs = []
        for n in (16, 32, 40):
            A = MatElliptic([n, n], gamma=4.0, exponent=2.0)
            u, u2 = A.create_exact_solution(0, cos_scale=cs)
            errs.append(np.abs(A.form_function(u)).max())
        assert errs[1] < 1e-6 * errs[0]
        assert errs[2] < 1e-9


def test_matmult_is_jacobian_of_function():
    # MatMult_Elliptic is the Gateaux derivative of FormFunction at the cached state
    rng = np.random.default_rng(0)
    A = MatElliptic([7, 6, 5], gamma=4.0, exponent=2.0)
    A.create_exact_solution(2)
    U = 0.5 + 0.1 * rng.standard_normal(A.g)
    V = rng.standard_normal(A.g)
    h = 1e-6
    Fp = A.form_function(U + h * V)
    Fm = A.form_function(U - h * V)
    A.form_function(U)
    J = A.mat_mult(V)
    assert np.abs((Fp - Fm) / (2 * h) - J).max() / np.abs(J).max() < 1e-6


def test_jacobian_matrix_is_fd_laplacian():
    A = MatElliptic([6, 6])
    A.create_exact_solution(1)
    A.form_function(np.zeros(A.g))
    P = A.form_jacobian_matrix()
    assert P.shape == (A.g, A.g)
    assert P.nnz == 5 * A.g - 2 * 4 * 2  # 5-point stencil minus links to Dirichlet nodes
    # symmetric positive definite for the constant-coefficient problem scaled by cell widths
    assert np.all(P.diagonal() > 0)
