"""The reference's own stokes.C + util.C (oracle/_ref/libstokesref.so, compiled unmodified against PETSc / FFTW / CppAD stand-ins)
against the numpy oracle on identical inputs - this is what pins oracle/stokes.py to the reference source - and, on a GPU,
against the CUDA shells through the C ABI."""
import numpy as np
import pytest

from oracle import ref
from oracle.stokes import StokesCtx, continuation_params
from conftest import rel_max

needs_ref = pytest.mark.skipif(not ref.stokes_available(), reason="oracle/_ref/libstokesref.so not built (needs /root/reference at build time)")

CASES = [([8, 6], 0, 2), ([8, 6], 1, 2), ([9, 7, 6], 1, 2), ([10, 10, 10], 0, 2), ([12, 11, 10], 1, 2), ([16, 16, 16], 1, 2), ([8, 6], 1, 1), ([7, 6, 5], 0, 1)]


def pair(dim, rheology, exact, exponent=3.0, eps=1e-2):
    R = ref.RefStokes(dim, rheology=rheology, hardness=1.0, exponent=exponent, regularization=eps, gamma0=1.0, exact=exact)
    O = StokesCtx(dim, rheology=rheology, hardness=1.0, exponent=exponent, regularization=eps, gamma0=1.0, exact=exact)
    U, U2 = O.create_exact_solution()
    return R, O, U, U2


@needs_ref
@pytest.mark.parametrize("dim,rheology,exact", CASES, ids=lambda v: str(v))
def test_numpy_oracle_equals_reference_source(dim, rheology, exact):
    R, O, U, U2 = pair(dim, rheology, exact)
    # StokesSetupDomain: DOF distribution (stokes.C:891) and the maps, seen through the scattered exact solution
    assert (R.m, R.g, R.gp, R.gv, R.dv) == (O.m, O.g, O.gp, O.gv, O.dv)
    # velocity part only: StokesExact2 leaves the pressure of a 3-D solution uninitialised in the reference (stokes.C:2001-2004)
    vo, _ = O.split(U)
    vr, _ = O.split(R.u)
    assert np.abs(vo - vr).max() < 1e-14 * max(np.abs(vo).max(), 1e-300)
    assert np.abs(O.dirichlet.reshape(-1) - R.dirichlet).max() < 1e-14
    assert np.abs(O.force - R.force).max() < 1e-12 * max(np.abs(R.force).max(), 1.0)
    xs = 0.3 * np.random.default_rng(1).standard_normal(O.g)
    assert rel_max(O.function(xs), R.function(xs)) < 1e-12
    assert rel_max(O.eta, R.eta) < 1e-13
    if rheology:
        assert rel_max(O.deta, R.deta) < 1e-13
    for j in range(O.d):
        assert rel_max(O.strain[j].reshape(-1), R.strain(j)) < 1e-12
    x = np.random.default_rng(0).standard_normal(O.g)
    v, p = O.split(x)
    assert rel_max(O.mat_mult(x), R.mat_mult(x)) < 1e-12
    assert rel_max(O.mat_mult_vv(v), R.mat_mult_vv(v)) < 1e-12
    assert rel_max(O.mat_mult_pv(v), R.mat_mult_pv(v)) < 1e-12
    assert rel_max(O.mat_mult_vp(p), R.mat_mult_vp(p)) < 1e-12
    assert rel_max(O.get_diagonal_schur(), R.get_diagonal_schur()) < 1e-13
    assert rel_max(O.mat_mult_schur(p, lambda rhs: rhs), R.mat_mult_schur_identity(p)) < 1e-12


@needs_ref
def test_pressure_reduce_order_and_pc_matrix():
    R, O, U, U2 = pair([12, 11, 10], 1, 2)
    pL = np.zeros(O.m)
    pL[O.int_nodes] = np.random.default_rng(3).standard_normal(O.gp)
    assert rel_max(O.pressure_reduce_order(pL.copy()), R.pressure_reduce_order(pL)) < 1e-12
    xs = 0.3 * np.random.default_rng(1).standard_normal(O.g)
    O.function(xs)
    R.function(xs)
    Po, Pr = O.pc_velocity_matrix(), R.pc_velocity_matrix()
    assert Po.shape == Pr.shape and abs(Po - Pr).max() < 1e-11 * abs(Pr).max()


@needs_ref
def test_continuation_and_exact_residual():
    # stokes.C:217-221 continuation parameters applied to the reference context; stokes.C:190-196 residual at the exact solution
    R, O, U, U2 = pair([10, 10, 10], 1, 2, exponent=3.0, eps=1e-4)
    xs = 0.3 * np.random.default_rng(5).standard_normal(O.g)
    for i in range(5):
        e, r = continuation_params(i, 4, 3.0, 1e-4)
        O.set_rheology(e, r)
        R.set_rheology(e, r)
        assert rel_max(O.function(xs), R.function(xs)) < 1e-12
    R0 = ref.RefStokes([20, 20, 20], rheology=0, exact=2)
    O0 = StokesCtx([20, 20, 20], rheology=0, exact=2)
    U0, _ = O0.create_exact_solution()
    assert np.abs(R0.function(U0)).max() < 2e-11  # "norm of residual" printed at stokes.C:196


@needs_ref
def test_golden_vectors_are_what_the_reference_source_produces():
    import os

    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name in ("stokes_8x6.npz", "stokes_9x7x6.npz"):
        z = np.load(os.path.join(G, name))
        R = ref.RefStokes([int(v) for v in z["dim"]], rheology=1, hardness=1.0, exponent=3.0, regularization=1e-2, gamma0=1.0, exact=2)
        assert rel_max(z["F"], R.function(z["xs"])) < 1e-12
        assert rel_max(z["y"], R.mat_mult(z["x"])) < 1e-12


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("dim,rheology", [([8, 6], 1), ([9, 7, 6], 1), ([16, 16, 16], 1), ([20, 20, 20], 0), ([32, 32, 32], 1)], ids=lambda v: str(v))
def test_cuda_shells_equal_reference_source(cuda, dim, rheology):
    import torch

    import spectral_petsc_b200 as sp

    R = ref.RefStokes(dim, rheology=rheology, hardness=1.0, exponent=3.0, regularization=1e-2, gamma0=1.0, exact=2)
    S = sp.Stokes(dim, rheology=rheology, hardness=1.0, exponent=3.0, regularization=1e-2, gamma0=1.0)
    assert (S.m, S.g, S.gp, S.gv, S.dv) == (R.m, R.g, R.gp, R.gv, R.dv)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    S.set_dirichlet(dev(R.dirichlet))
    S.set_force(dev(R.force))
    xs = 0.3 * np.random.default_rng(1).standard_normal(R.g)
    x = np.random.default_rng(0).standard_normal(R.g)
    assert rel_max(S.function(dev(xs)).cpu().numpy(), R.function(xs)) < 1e-12
    assert rel_max(S.get_state(0).cpu().numpy(), R.eta) < 1e-13
    assert rel_max(S.mat_mult(dev(x)).cpu().numpy(), R.mat_mult(x)) < 1e-12
