"""The reference's own chebyshev.c (oracle/_ref/libchebref.so, built by oracle/Makefile from /root/reference against FFTW /
PETSc stand-ins) against (i) the numpy oracle, (ii) the differentiation matrix the CUDA kernels apply, (iii) the reference's
analytic known answers, (iv) its error behaviour - and, on a GPU, against the CUDA ChebMult through the C ABI."""
import ctypes

import numpy as np
import pytest

from oracle import ref
from oracle.chebyshev import ChebCtx, cheb_mult
from conftest import rel_max

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libchebref.so not built (needs /root/reference at build time)")

SHAPES = [(1, 0, [5]), (1, 0, [128]), (2, 0, [8, 6]), (2, 1, [8, 6]), (3, 0, [8, 7, 6]), (3, 1, [8, 7, 6]), (3, 2, [8, 7, 6]),
          (3, 0, [16, 16, 16]), (3, 1, [33, 9, 4]), (3, 2, [6, 5, 128]), (4, 2, [5, 4, 20, 3]), (5, 3, [4, 3, 5, 12, 3]),
          (4, 0, [32, 4, 4, 3]), (4, 2, [4, 4, 32, 3])]  # the last two: velocity layout with trailing component axis (stokes.C:284-289)


@needs_ref
@pytest.mark.parametrize("rank,tr,dims", SHAPES, ids=lambda v: str(v))
def test_numpy_oracle_equals_reference_source(rank, tr, dims):
    x = np.random.default_rng(0).standard_normal(int(np.prod(dims)))
    yr = ref.cheb_mult(rank, tr, dims, x)
    yo = cheb_mult(ChebCtx(rank, tr, dims), x)
    assert rel_max(yo, yr) < 2e-13 * max(1.0, (dims[tr] / 32.0) ** 2)


@needs_ref
@pytest.mark.parametrize("P", [5, 12, 16, 20, 32, 64, 128, 129])
def test_product_matrix_equals_reference_source(P):
    """The P x P matrix the CUDA kernels apply (host function of the C ABI, no GPU needed) against the reference's ChebMult."""
    import spectral_petsc_b200 as sp

    D = sp.cheb_matrix(P)
    x = np.random.default_rng(1).standard_normal(P)
    yr = ref.cheb_mult(1, 0, [P], x)
    assert rel_max(D @ x, yr) < 1e-12
    assert rel_max(ref.chebd1_mult(x), yr) < 1e-13  # the 1-D and the guru code paths of the reference agree


@needs_ref
def test_reference_known_answers_K1_K2():
    # cheb.c:68-70,95-103: u = exp(cos(i pi/(m-1))), expect Du = u; printed error norm 1.029e-02 at the default m1 = 5
    for m, bound in ((5, 2e-2), (8, 2e-5), (16, 1e-12)):
        xi = np.cos(np.arange(m) * np.pi / (m - 1))
        du = ref.cheb_mult(1, 0, [m], np.exp(xi))
        assert np.abs(du - np.exp(xi)).max() < bound
    dims = [8, 7, 6]
    idx = np.indices(dims).reshape(3, -1)
    X = [np.cos(idx[j] * np.pi / (dims[j] - 1)) for j in range(3)]
    u = np.exp(X[0]) + np.exp(X[1]) + np.exp(X[2])
    for tr, bound in ((0, 2e-5), (1, 3e-4), (2, 3e-3)):
        assert np.abs(ref.cheb_mult(3, tr, dims, u) - np.exp(X[tr])).max() < bound
    z = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "cheb_8x7x6.npz"))
    for tr in range(3):  # the committed golden vectors are what the reference source produces
        assert rel_max(z["du%d" % tr], ref.cheb_mult(3, tr, dims, z["u"])) < 1e-13
        assert rel_max(z["drnd%d" % tr], ref.cheb_mult(3, tr, dims, z["rnd"])) < 1e-13


@needs_ref
def test_reference_error_behaviour_matches_c_abi():
    """chebyshev.c:98,106,122 raise PETSC_ERR_USER (83); the C ABI returns the same code for the same inputs."""
    import spectral_petsc_b200 as sp

    L = sp.lib()
    h = ctypes.c_void_p()
    for rank, tr, dims, n in ((2, 2, [4, 4], 16), (2, 0, [4, 4], 15), (1, 0, [1], 1)):
        with pytest.raises(ref.RefError) as ei:
            x = np.zeros(n)
            y = np.zeros(n)
            arr = (ctypes.c_int * len(dims))(*dims)
            rc = ref.lib().ref_cheb_mult(rank, tr, arr, n, x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p))
            if rc:
                raise ref.RefError(rc, ref.lib().ref_last_error().decode())
        assert ei.value.code == 83
        arr = (ctypes.c_int * len(dims))(*dims)
        assert L.sb200_cheb_create(rank, tr, arr, ctypes.c_longlong(n), ctypes.byref(h)) == 83


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("rank,tr,dims", [(3, 0, [16, 16, 16]), (3, 2, [8, 7, 6]), (3, 1, [32, 32, 32]), (1, 0, [128]), (4, 2, [4, 4, 32, 3]), (3, 0, [64, 16, 8])],
                         ids=lambda v: str(v))
def test_cuda_chebmult_equals_reference_source(cuda, rank, tr, dims):
    import torch

    import spectral_petsc_b200 as sp

    x = np.random.default_rng(0).standard_normal(int(np.prod(dims)))
    y = sp.Cheb(rank, tr, dims).mult(torch.from_numpy(x).to(cuda)).cpu().numpy()
    assert rel_max(y, ref.cheb_mult(rank, tr, dims, x)) < 1e-12
