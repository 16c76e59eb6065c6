"""Solver orchestration (spectral_petsc_b200/solvers.py: StokesPCApply0..3, the inner left-preconditioned GMRES solves, the
outer FGMRES with the constant-pressure null space, the Newton skeleton) driven over the CPU oracle: BASELINE config 4
(./stokes -exact 2 -schur_ksp_max_it 3 -vel_ksp_max_it 4 -svel_ksp_type preonly -ksp_type fgmres -ksp_rtol 1e-10, linear
viscosity; README:47) converges to the manufactured solution.  The same code drives the CUDA shells in tests/test_gpu_solvers.py."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle.elliptic import MatElliptic
from oracle.fgmres import fgmres
from oracle.stokes import StokesCtx
from spectral_petsc_b200 import solvers


def np_krylov(op, b, pc, rtol, maxits, restart):
    x, its, hist, reason = fgmres(op, b, M=pc, restart=restart, rtol=rtol, maxits=maxits)
    return x, its, reason


def stokes_problem(dim):
    O = StokesCtx(dim, rheology=0, exact=2)
    U, _ = O.create_exact_solution()
    F0 = O.function(np.zeros(O.g))
    lu = spla.splu(O.pc_velocity_matrix().tocsc())  # stand-in for -vel_pc_type hypre on MatVVPC
    return O, U, F0, lu


def test_split_merge_roundtrip():
    x = np.arange(40.0)
    v, p = solvers.split(x, 3)
    assert v.size == 30 and p.size == 10 and p[1] == 7.0 and v[3] == 4.0
    assert np.array_equal(solvers.merge(v, p, 3), x)
    y = solvers.remove_constant_pressure(x, 3)
    assert abs(solvers.split(y, 3)[1].mean()) < 1e-13 and np.array_equal(solvers.split(y, 3)[0], v)


@pytest.mark.parametrize("saddle", [0, 1, 2, 3])
def test_config4_stokes_linear_converges_to_exact_solution(saddle):
    dim = [12, 12, 12]
    O, U, F0, lu = stokes_problem(dim)
    pc = solvers.StokesSaddlePC(O, 3, np_krylov, lu.solve, saddle_type=saddle)
    dx, its, reason = solvers.solve_stokes_linear(O, 3, np_krylov, pc, -F0, rtol=1e-10, maxits=200)
    assert reason == 2, (its, reason)
    assert its <= (40 if saddle != 2 else 90)
    v, p = solvers.split(dx, 3)
    ve, pe = solvers.split(U, 3)
    # "Norm of error" (stokes.C:226-233): spectral accuracy of sin/cos on 12 nodes
    assert np.abs(v - ve).max() < 1e-6
    assert np.abs((p - p.mean()) - (pe - pe.mean())).max() < 1e-4
    assert pc.inner_its["velocity"] > 0 and pc.inner_its["schur"] > 0
    # the residual of the full operator really dropped by 1e-10
    assert np.linalg.norm(O.mat_mult(dx) + F0) <= 2e-10 * np.linalg.norm(F0)


def test_newton_elliptic_nonlinear():
    # tests.sh: ./elliptic -dim n,n -exact 0 -cos_scale s -gamma 4 -ksp_rtol 1e-12 -snes_rtol 1e-12
    O = MatElliptic([16, 16], gamma=4.0, exponent=2.0)
    u, _ = O.create_exact_solution(0, cos_scale=1.0)

    def solve_jacobian(rhs):
        lu = spla.splu(O.form_jacobian_matrix().tocsc())
        x, its, _ = np_krylov(O.mat_mult, rhs, lu.solve, 1e-12, 200, 30)
        return x, its

    x, its, kits, hist = solvers.newton(O.form_function, solve_jacobian, np.zeros(O.g), rtol=1e-12)
    assert its <= 12 and hist[-1] <= 1e-12 * hist[0]
    assert np.abs(x - u).max() < 1e-9  # "Norm of error"
    assert all(h2 < h1 for h1, h2 in zip(hist, hist[1:]))


def test_config5_power_law_continuation_small():
    """./stokes -exact 2 -cont 4 -rheology 1 -eps 1e-4 -exponent 3 ... (README:55) at a small extent: five SNES solves, the first
    one linear, each started from the previous solution; every step converges and the viscosity contrast grows."""
    from oracle.stokes import continuation_params as oracle_params

    dim = [8, 8, 8]
    O = StokesCtx(dim, rheology=1, hardness=1.0, exponent=3.0, regularization=1e-4, gamma0=1.0, exact=2)
    O.create_exact_solution()
    for i in range(5):  # the product-side formula equals the oracle's restatement of stokes.C:217-219
        assert solvers.continuation_params(i, 4, 3.0, 1e-4) == pytest.approx(oracle_params(i, 4, 3.0, 1e-4), rel=1e-15)

    def make_pc():
        lu = spla.splu(O.pc_velocity_matrix().tocsc())
        return solvers.StokesSaddlePC(O, 3, np_krylov, lu.solve, saddle_type=0)

    x, log = solvers.solve_stokes_continuation(O.function, O.set_rheology, make_pc, O, 3, np_krylov, np.zeros(O.g), 3.0, 1e-4, cont=4,
                                               ksp_rtol=1e-6, snes_rtol=1e-8, ksp_maxits=300)
    assert [s["step"] for s in log] == [0, 1, 2, 3, 4]
    assert log[0]["exponent"] == 1.0 and log[-1]["exponent"] == pytest.approx(3.0) and log[-1]["regularization"] == pytest.approx(1e-4)
    assert log[0]["snes_its"] <= 2  # the linear problem
    for s in log:
        assert s["fnorm"][-1] <= 1e-8 * s["fnorm"][0] * 1.01 or s["fnorm"][-1] < 1e-10, s
        assert s["snes_its"] <= 30
    assert O.max_eta / O.min_eta > 3.0  # shear-thinning state at the end (stokes.C:731-734 prints these)
