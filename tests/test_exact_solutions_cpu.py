"""The manufactured solutions of the drivers (sb200_elliptic_exact_solution = CreateExactSolution elliptic.C:594-677,
sb200_stokes_exact_solution = StokesCreateExactSolution + StokesExact0..3 + StokesDirichlet, stokes.C:942-1003,1948-2050) are
host functions of the C-ABI library (set-up work, no device): compared here with the oracle, which tests/test_oracle_ref_*.py
pin against the reference source."""
import numpy as np
import pytest

import spectral_petsc_b200 as sp
from oracle.elliptic import MatElliptic
from oracle.stokes import StokesCtx


@pytest.mark.parametrize("dim", [[8, 6], [16, 16, 16], [12] * 5, [5, 4, 3, 6], [3, 3]], ids=str)
@pytest.mark.parametrize("exact,cos_scale,gamma,exponent", [(0, 3.0, 4.0, 2.0), (0, 2.8, 4.0, 2.0), (0, 1.0, 0.0, 0.0), (1, 0.0, 0.0, 2.0), (2, 0.0, 4.0, 3.0)])
def test_elliptic_exact_solution_equals_oracle(dim, exact, cos_scale, gamma, exponent):
    O = MatElliptic(dim, gamma=gamma, exponent=exponent)
    with np.errstate(all="ignore"):
        u, u2 = O.create_exact_solution(exact, cos_scale=cos_scale)
    gu, gu2, gd = sp.elliptic_exact_solution(dim, exact, cos_scale, gamma, exponent)
    # same libm, same operation order: the only freedom is the compiler's contraction of a*b+c
    for a, b in ((gu, u), (gu2, u2), (gd, O.dirichlet)):
        assert a.shape == b.shape
        assert np.allclose(a, b, rtol=1e-14, atol=1e-14 * max(1.0, np.abs(b).max()), equal_nan=True)


def test_elliptic_exact_solution_errors():
    with pytest.raises(sp.SB200Error) as ei:
        sp.elliptic_exact_solution([8, 6], 3)  # "Choose an exact solution." (elliptic.C:657)
    assert ei.value.code == 83 and "Choose an exact solution" in str(ei.value)
    with pytest.raises(sp.SB200Error):
        sp.elliptic_exact_solution([8, 2], 1)


@pytest.mark.parametrize("dim", [[8, 6], [7, 6, 5], [20, 20, 20]], ids=str)
@pytest.mark.parametrize("exact", [0, 1, 2])
def test_stokes_exact_solution_equals_oracle(dim, exact):
    O = StokesCtx(dim, exact=exact)
    U, U2 = O.create_exact_solution()
    gU, gU2, gd = sp.stokes_exact_solution(dim, exact)
    assert np.allclose(gU, U, rtol=1e-15, atol=1e-15)
    assert np.allclose(gU2, U2, rtol=1e-15, atol=1e-14)
    assert np.allclose(gd, O.dirichlet.reshape(-1), rtol=1e-15, atol=1e-15)


def test_stokes_exact3_is_two_dimensional_shear():
    U, U2, dr = sp.stokes_exact_solution([6, 5], 3)  # StokesExact3 (stokes.C:2016-2034): u = y + 1, v = p = 0, no forcing
    y = np.cos(np.arange(1, 4) * np.pi / 4)
    assert np.allclose(U.reshape(4, 3, 3)[:, :, 0], (y + 1.0)[None, :]) and not U.reshape(-1, 3)[:, 1:].any() and not U2.any()
    with pytest.raises(sp.SB200Error) as ei:
        sp.stokes_exact_solution([6, 5, 4], 3)
    assert ei.value.code == 83
    with pytest.raises(sp.SB200Error) as ei:
        sp.stokes_exact_solution([6, 5], 4)
    assert ei.value.code == 56  # PETSC_ERR_SUP (stokes.C:452)
