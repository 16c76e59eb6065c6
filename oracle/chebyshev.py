"""Oracle restatement of chebyshev.c (test infrastructure, see oracle/__init__.py).

Follows /root/reference/chebyshev.c:
  MatCreateCheb  :89-138   plan / context (rank, tr, row-major strides, work array)
  ChebMult       :142-199  DCT-I -> *k + endpoint sums -> DST-I -> 1/(2n sin) scaling
  ChebD1Mult     :37-71    the 1-D variant (same arithmetic, different loop fusion)

FFTW_REDFT00 / FFTW_RODFT00 (unnormalised) are ``scipy.fft.dct/dst(type=1)``.
"""
import math

import numpy as np
import scipy.fft

PI = 3.14159265358979323846  # chebyshev.h:10


class ChebError(ValueError):
    """Stands in for SETERRQ(PETSC_ERR_USER, ...) in chebyshev.c:98,106,122."""


class ChebCtx:
    """MatCreateCheb (chebyshev.c:89-138).

    ``dims`` is the row-major extent list (last axis fastest, :107-120) and
    ``tr`` the transformed axis.  ``n_total`` is the Vec length it must match.
    """

    def __init__(self, rank, tr, dims, n_total=None, workers=1):
        dims = [int(x) for x in dims[:rank]]
        n = int(np.prod(dims)) if n_total is None else int(n_total)
        if n < 2:
            raise ChebError("n = %d but must be >= 2" % n)  # :98
        if not (0 <= tr < rank):
            raise ChebError("tdim out of range")  # :106
        stride = int(np.prod(dims))
        if n != stride:
            raise ChebError("dimensions do not agree: n = %d but stride = %d" % (n, stride))  # :122
        self.rank, self.tr, self.dims, self.N = rank, tr, dims, n
        self.workers = workers

    def mult(self, x):
        return cheb_mult(self, x)


def cheb_mult(c, x):
    """ChebMult (chebyshev.c:142-199).  x: flat array of length N (preserved); returns y."""
    x = np.asarray(x, dtype=np.float64)
    if x.size != c.N:
        raise ChebError("size mismatch")
    X = x.reshape(c.dims)
    tr = c.tr
    n = c.dims[tr] - 1  # :154 nodes are numbered [0..n]
    N = float(n)
    # :157 forward REDFT00 along the transformed axis, all other axes are "howmany" loops
    work = scipy.fft.dct(X, type=1, axis=tr, workers=c.workers)
    work = np.moveaxis(work, tr, 0)  # view; line index first
    y = np.empty_like(work)
    y0 = np.zeros(work.shape[1:])
    yn = np.zeros(work.shape[1:])
    s = 1.0
    # :168-175  work[i] *= i; endpoint sums accumulated in index order
    for i in range(1, n):
        I = float(i)
        work[i] *= I
        y0 += I * work[i]
        yn += s * I * work[i]
        s = -s
    y[0] = 0.5 * work[n] * N + y0 / n  # :176
    y[n] = yn / N + 0.5 * s * N * work[n]  # :177
    if n > 1:
        # :181 backward RODFT00 of length n-1 on the interior coefficients
        y[1:n] = scipy.fft.dst(work[1:n], type=1, axis=0, workers=c.workers)
        pin = PI / N  # :183
        idx = np.arange(1, n, dtype=np.float64)
        # :190   y /= 2 * n * sqrt(1 - cos(i pi/n)^2)
        scale = (2 * n) * np.sqrt(1.0 - np.cos(idx * pin) ** 2)
        y[1:n] /= scale.reshape((-1,) + (1,) * (y.ndim - 1))
    return np.ascontiguousarray(np.moveaxis(y, 0, tr)).reshape(-1)


def cheb_d1_mult(x):
    """ChebD1Mult (chebyshev.c:37-71), 1-D operator on n+1 points."""
    x = np.asarray(x, dtype=np.float64)
    if x.size < 2:
        raise ChebError("n = %d but must be >= 2" % x.size)  # :18
    n = x.size - 1
    work = scipy.fft.dct(x, type=1)
    for i in range(1, n):
        work[i] *= float(i)  # :51
    y = np.zeros_like(x)
    if n > 1:
        y[1:n] = scipy.fft.dst(work[1:n], type=1)  # :53
    N = float(n)
    pin = PI / N
    s = 1.0
    y0 = 0.0
    yn = 0.0
    for i in range(1, n):  # :60-66
        I = float(i)
        y[i] /= 2.0 * n * math.sqrt(1.0 - math.cos(I * pin) ** 2)
        y0 += I * work[i]
        yn += s * I * work[i]
        s = -s
    y[0] = 0.5 * work[n] * N + y0 / n  # :67
    y[n] = yn / N + 0.5 * s * N * work[n]  # :68
    return y


def cgl_nodes(npts):
    """Collocation nodes as the drivers build them: cos(i*pi/(dim-1)) (elliptic.C:279, stokes.C:296)."""
    return np.cos(np.arange(npts) * math.pi / (npts - 1))


def dense_cgl_matrix(npts):
    """Classical CGL differentiation matrix (textbook formula, float64).

    Not part of the reference - used by tests to confirm that the restated FFT
    path is the CGL derivative and by nothing else.
    """
    n = npts - 1
    x = np.cos(np.pi * np.arange(npts) / n)
    c = np.ones(npts)
    c[0] = c[n] = 2.0
    c *= (-1.0) ** np.arange(npts)
    X = np.tile(x, (npts, 1)).T
    dX = X - X.T
    D = np.outer(c, 1.0 / c) / (dX + np.eye(npts))
    D -= np.diag(D.sum(axis=1))
    return D
