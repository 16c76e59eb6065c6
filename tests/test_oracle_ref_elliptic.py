"""The reference's own elliptic.C (oracle/_ref/libellipticref.so: MatCreate_Elliptic, SetupBC, CreateExactSolution, FormFunction,
MatMult_Elliptic, FormJacobian compiled unmodified against PETSc / FFTW stand-ins) against the numpy oracle on identical
inputs - this is what pins oracle/elliptic.py to the reference source - and, on a GPU, against the CUDA shells through the C ABI."""
import numpy as np
import pytest

from oracle import ref
from oracle.elliptic import MatElliptic
from conftest import rel_max

needs_ref = pytest.mark.skipif(not ref.elliptic_available(), reason="oracle/_ref/libellipticref.so not built (needs /root/reference at build time)")

CASES = [([8, 6], 0.0, 2.0, 2), ([8, 6], 4.0, 2.0, 2), ([7, 6, 5], 4.0, 2.0, 2), ([16, 16, 16], 0.0, 2.0, 2), ([16, 16, 16], 4.0, 2.0, 2),
         ([12, 12, 12], 4.0, 3.0, 1), ([6, 5, 4, 5], 1.5, 2.0, 2), ([6] * 5, 4.0, 2.0, 2), ([32, 20], 4.0, 2.0, 0)]


@needs_ref
@pytest.mark.parametrize("dim,gamma,exponent,exact", CASES, ids=lambda v: str(v))
def test_numpy_oracle_equals_reference_source(dim, gamma, exponent, exact):
    R = ref.RefElliptic(dim, gamma=gamma, exponent=exponent, exact=exact, cos_scale=1.0)
    O = MatElliptic(dim, gamma=gamma, exponent=exponent)
    u, u2 = O.create_exact_solution(exact, cos_scale=1.0)
    # SetupBC: DOF counts (printed at elliptic.C:424) and the walk order through the scattered exact solution
    assert (R.m, R.g, R.nd) == (O.m, O.g, O.nd)
    assert rel_max(u, R.u) < 1e-14 and rel_max(u2, R.u2) < 1e-13
    assert rel_max(O.dirichlet, R.dirichlet) < 1e-14 or np.abs(R.dirichlet).max() == 0
    assert rel_max(O.b, R.b) < 1e-13
    Us = 0.1 * np.random.default_rng(1).standard_normal(O.g)
    U = np.random.default_rng(0).standard_normal(O.g)
    assert rel_max(O.form_function(Us), R.form_function(Us)) < 1e-12
    assert rel_max(O.eta, R.eta) < 1e-14
    if gamma:
        assert rel_max(O.deta, R.deta) < 1e-14
    for k in range(O.d):
        assert rel_max(O.gradu[k], R.gradu(k)) < 1e-12
    assert rel_max(O.mat_mult(U), R.mat_mult(U)) < 1e-12
    # FormJacobian: the finite-difference preconditioning matrix
    Jo, Jr = O.form_jacobian_matrix(), R.jacobian()
    assert Jo.shape == Jr.shape and abs(Jo - Jr).max() < 1e-11 * abs(Jr).max()


@needs_ref
def test_reference_exact2_residual_K3():
    # "Norm of exact residual" (elliptic.C:187-201): F(u_exact) ~ 1e-12 for the linear problem
    R = ref.RefElliptic([16, 16, 16], gamma=0.0, exact=2)
    assert np.abs(R.form_function(R.u)).max() < 5e-11
    R = ref.RefElliptic([12] * 5, gamma=0.0, exact=2)
    assert (R.m, R.g) == (248832, 100000)
    assert np.abs(R.form_function(R.u)).max() < 5e-11


@needs_ref
def test_golden_vectors_are_what_the_reference_source_produces():
    import os

    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name in ("elliptic_8x6.npz", "elliptic_7x6x5.npz", "elliptic_16x16x16.npz"):
        z = np.load(os.path.join(G, name))
        R = ref.RefElliptic([int(v) for v in z["dim"]], gamma=4.0, exponent=2.0, exact=2)
        assert rel_max(z["F"], R.form_function(z["Us"])) < 1e-12
        assert rel_max(z["V"], R.mat_mult(z["U"])) < 1e-12
        assert rel_max(z["dirichlet"], R.dirichlet) < 1e-14


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("dim,gamma", [([16, 16, 16], 4.0), ([8, 6], 4.0), ([32, 32, 32], 4.0), ([12] * 4, 0.0), ([64, 64], 4.0)], ids=lambda v: str(v))
def test_cuda_shells_equal_reference_source(cuda, dim, gamma):
    import torch

    import spectral_petsc_b200 as sp

    R = ref.RefElliptic(dim, gamma=gamma, exponent=2.0, exact=2)
    G = sp.Elliptic(dim, gamma=gamma, exponent=2.0)
    assert (G.m, G.g, G.nd) == (R.m, R.g, R.nd)
    G.set_dirichlet(torch.from_numpy(R.dirichlet).to(cuda))
    G.set_rhs(torch.from_numpy(R.b).to(cuda))
    Us = 0.1 * np.random.default_rng(1).standard_normal(R.g)
    U = np.random.default_rng(0).standard_normal(R.g)
    assert rel_max(G.form_function(torch.from_numpy(Us).to(cuda)).cpu().numpy(), R.form_function(Us)) < 1e-12
    assert rel_max(G.get_state(0).cpu().numpy(), R.eta) < 1e-14
    assert rel_max(G.mat_mult(torch.from_numpy(U).to(cuda)).cpu().numpy(), R.mat_mult(U)) < 1e-12
