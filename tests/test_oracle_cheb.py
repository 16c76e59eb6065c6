"""Pins the oracle's ChebMult restatement against the reference's own known-answer tests
(cheb.c) and against the textbook CGL matrix."""
import numpy as np
import pytest

from oracle.chebyshev import PI, ChebCtx, ChebError, cheb_d1_mult, cheb_mult, dense_cgl_matrix
from conftest import rel_max


def test_K1_cheb_c_1d_exp():
    # cheb.c:68-70,95-103: u = exp(cos(i pi/(m1-1))), Du = u; default m1 = 5 prints 1.029e-02
    for m1, expect in ((5, 1.0293308609854e-02), (9, 3.9095e-07), (17, 3e-14)):
        u = np.exp(np.cos(np.arange(m1) * PI / (m1 - 1)))
        err = np.abs(cheb_d1_mult(u) - u).max()
        if m1 == 17:
            assert err < 1e-13
        else:
            assert err == pytest.approx(expect, rel=1e-3)
        assert np.array_equal(cheb_d1_mult(u), cheb_mult(ChebCtx(1, 0, [m1]), u))


def test_K2_cheb_c_3d_exp():
    # cheb.c:77-91,105-112 with (m,n,p) = (8,7,6)
    m, n, p = 8, 7, 6
    x, y, z = (np.cos(np.arange(k) * PI / (k - 1)) for k in (m, n, p))
    a = np.exp(x)[:, None, None] + np.exp(y)[None, :, None] + np.exp(z)[None, None, :]
    exact = [np.broadcast_to(np.exp(x)[:, None, None], a.shape), np.broadcast_to(np.exp(y)[None, :, None], a.shape),
             np.broadcast_to(np.exp(z)[None, None, :], a.shape)]
    expect = [6.245e-06, 8.72e-05, 1.04e-03]
    for d in range(3):
        err = np.abs(cheb_mult(ChebCtx(3, d, [m, n, p]), a.ravel()) - exact[d].ravel()).max()
        assert err == pytest.approx(expect[d], rel=5e-3)


@pytest.mark.parametrize("P", [2, 3, 5, 12, 16, 20, 33, 128, 129])
def test_fft_path_is_cgl_matrix(P):
    rng = np.random.default_rng(P)
    u = rng.standard_normal(P)
    D = dense_cgl_matrix(P)
    assert rel_max(cheb_d1_mult(u), D @ u) < 1e-12


def test_nd_axes_and_vector_layout():
    # stokes.C:284-289: rank d+1 with a trailing component axis of extent d
    rng = np.random.default_rng(1)
    dims = [6, 5, 7, 3]
    x = rng.standard_normal(dims)
    for tr in range(3):
        y = cheb_mult(ChebCtx(4, tr, dims), x.ravel()).reshape(dims)
        D = dense_cgl_matrix(dims[tr])
        ref = np.moveaxis(np.tensordot(D, np.moveaxis(x, tr, 0), axes=(1, 0)), 0, tr)
        assert rel_max(y, ref) < 1e-13


def test_errors_like_reference():
    with pytest.raises(ChebError):
        ChebCtx(1, 0, [1])  # chebyshev.c:98
    with pytest.raises(ChebError):
        ChebCtx(2, 2, [4, 4])  # :106
    with pytest.raises(ChebError):
        ChebCtx(2, 0, [4, 4], n_total=15)  # :122
