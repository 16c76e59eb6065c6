"""Pins the oracle's stokes.C restatement against the reference's self-checks: DOF counts
(stokes.C:891), exact-solution residual (stokes.C:190-196), the constant-pressure null space that
MatNullSpaceTest asserts (stokes.C:206-212), util.C's polyInterp self test, the rheology laws."""
import math

import numpy as np
import pytest

from oracle.stokes import StokesCtx, continuation_params, poly_interp, rheology_power


def test_K8_dof_counts():
    S = StokesCtx([20, 20, 20])
    assert (S.g, S.gp, S.gv, S.dv, S.m) == (23328, 5832, 17496, 6504, 8000)


def test_K5_exact2_residual_3d():
    S = StokesCtx([20, 20, 20], exact=2)
    U, U2 = S.create_exact_solution()
    r = S.function(U)
    assert np.abs(U).max() == pytest.approx(0.99137, rel=1e-4)
    assert np.abs(r).max() < 2e-11  # 3.2e-12 at 20^3
    assert (S.min_eta, S.max_eta) == (1.0, 1.0)


def test_K5_exact_residual_converges_2d():
    errs = []
    for n in (8, 12, 16, 24):
        S = StokesCtx([n, n], exact=1)
        U, _ = S.create_exact_solution()
        errs.append(np.abs(S.function(U)).max())
    assert errs[3] < 1e-8 * errs[0]


@pytest.mark.parametrize("dim", [[8, 6], [9, 7, 6], [16, 16, 16]])
def test_K6_constant_pressure_null_space(dim):
    S = StokesCtx(dim, rheology=1, exponent=3.0, regularization=1e-2, exact=2)
    U, _ = S.create_exact_solution()
    S.function(U)  # variable viscosity state
    ns = S.merge(np.zeros(S.gv), np.ones(S.gp))
    assert np.abs(S.mat_mult(ns)).max() < 1e-12


def test_K7_polyinterp_util_main():
    # util.C:155-171: cos on nodes 1..order, evaluated at 1.43 and 3.1
    for order in range(2, 20):
        x = 1.0 + np.arange(order)
        f0, f1 = poly_interp(order, x, np.cos(x), 1.43, 3.1)
        if order >= 12:
            assert abs(f0 - math.cos(1.43)) < 6e-3 and abs(f1 - math.cos(3.1)) < 1e-4
    # exactness on polynomials of degree < n, including extrapolation
    x = np.cos(np.arange(1, 9) * math.pi / 9)
    pfun = lambda t: 3 * t ** 7 - t ** 4 + 2 * t - 5
    f0, f1 = poly_interp(8, x, pfun(x), 1.0, -1.0)
    assert abs(f0 - pfun(1.0)) < 1e-12 and abs(f1 - pfun(-1.0)) < 1e-12


def test_pressure_reduce_order_reproduces_low_degree_pressure():
    # a pressure of degree <= P-3 per axis is reproduced exactly at the boundary nodes
    S = StokesCtx([9, 8, 7])
    c = S.coord
    p_exact = (1 + c[:, 0] ** 3) * (2 - c[:, 1] ** 2) * (1 + 0.5 * c[:, 2] ** 4)
    pL = np.zeros(S.m)
    pL[S.int_nodes] = p_exact[S.int_nodes]
    S.pressure_reduce_order(pL)
    assert np.abs(pL - p_exact).max() < 1e-12


def test_rheology_and_continuation():
    g = np.array([0.0, 0.5, 2.0])
    eta, deta = rheology_power(g, 2.0, 3.0, 1e-4, 1.5)
    p = (1 - 3.0) / 6.0
    assert np.allclose(eta, 2.0 * (1e-4 + g / 1.5) ** p)
    assert np.allclose(deta, 2.0 * p / 1.5 * (1e-4 + g / 1.5) ** (p - 1))
    e0, r0 = continuation_params(0, 4, 3.0, 1e-4)
    e4, r4 = continuation_params(4, 4, 3.0, 1e-4)
    assert (e0, r0) == (1.0, 1.0) and e4 == pytest.approx(3.0) and r4 == pytest.approx(1e-4)
    e2, r2 = continuation_params(2, 4, 3.0, 1e-4)
    assert e2 == pytest.approx(1 + 0.5 ** 0.8 * 2) and r2 == pytest.approx(1e-2)


def test_vv_is_jacobian_of_viscous_residual():
    rng = np.random.default_rng(0)
    S = StokesCtx([7, 6, 5], rheology=1, exponent=3.0, regularization=1e-1, exact=2)
    S.create_exact_solution()
    x = 0.3 * rng.standard_normal(S.g)
    dx = S.merge(rng.standard_normal(S.gv), np.zeros(S.gp))
    h = 1e-6
    Fp = S.function(x + h * dx)
    Fm = S.function(x - h * dx)
    S.function(x)
    J = S.mat_mult(dx)
    assert np.abs((Fp - Fm) / (2 * h) - J).max() / np.abs(J).max() < 1e-6


@pytest.mark.parametrize("dim", [[9, 7, 6], [12, 10]], ids=str)
def test_folded_form_of_the_block_operator(dim):
    """The identity behind the CUDA path's two opt-ins (sb200_stokes_set_trace_divergence / _set_fold_pressure): with the gradient
    G_j = D_j w of the padded velocity, PV v = trace(G) and VV v + VP p = -sum_j D_j (V_j - p_ext e_j), p_ext the padded and
    boundary-extrapolated pressure - i.e. StokesMatMult from ONE gradient and ONE divergence (18 scalar derivatives, not 24)."""
    O = StokesCtx(dim, rheology=1, exponent=2.0, regularization=0.5, exact=2)
    U, _ = O.create_exact_solution()
    O.function(U)
    d = O.d
    x = np.random.default_rng(0).standard_normal(O.g)
    y0 = O.mat_mult(x)
    v, p = O.split(x)
    xL = O.vel_local(v, False)
    G = [O.dvel(i, xL) for i in range(d)]
    div = np.zeros(O.m)
    for i in range(d):
        div = div + G[i][:, i]
    st = [[0.5 * (G[j][:, k] + G[k][:, j]) for k in range(d)] for j in range(d)]
    z = np.zeros(O.m)
    for j in range(d):
        for k in range(d):
            z = z + st[j][k] * O.strain[j][:, k]
    pL = np.zeros(O.m)
    pL[O.int_nodes] = p
    O.pressure_reduce_order(pL)
    yL = np.zeros((O.m, d))
    for j in range(d):
        Vj = np.empty((O.m, d))
        for k in range(d):
            Vj[:, k] = O.eta * st[j][k] + O.deta * O.strain[j][:, k] * z - (pL if j == k else 0.0)
        yL = yL + (-1.0) * O.dvel(j, Vj)
    y1 = O.merge(yL[O.int_nodes].reshape(-1), div[O.int_nodes])
    assert np.array_equal(O.split(y1)[1], O.split(y0)[1])          # the divergence rows: the same numbers, bit for bit
    assert np.abs(y1 - y0).max() <= 1e-14 * np.abs(y0).max()       # the velocity rows: one rounding of the sum instead of two
