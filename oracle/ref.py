"""ctypes access to oracle/_ref/libchebref.so: the REFERENCE's own chebyshev.c (MatCreateCheb / ChebMult / ChebDestroy,
MatCreateChebD1 / ChebD1Mult), compiled unmodified from /root/reference against minimal FFTW / PETSc stand-ins
(oracle/ref_stubs/, recipe oracle/Makefile).  Test infrastructure: it pins the numpy restatement (oracle/chebyshev.py) and
the product's differentiation matrix against the reference source itself.  FFTW's two r2r kinds are the stand-in's O(n^2)
implementation of their documented definitions, so what is pinned is everything chebyshev.c does around them (strides,
odometer loops, k-scaling, end-point sums, the 1/(2n sin) division)."""
import ctypes
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libchebref.so")
_lib = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(_PATH)
        L.ref_last_error.restype = ctypes.c_char_p
        _lib = L
    return _lib


class RefError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("reference error %d: %s" % (code, msg))
        self.code = code


def cheb_mult(rank, tr, dims, x):
    """The reference's ChebMult on a row-major array with extents dims[:rank] (chebyshev.c:142-199)."""
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    y = np.zeros_like(x)
    arr = (ctypes.c_int * len(dims))(*[int(v) for v in dims])
    rc = lib().ref_cheb_mult(ctypes.c_int(rank), ctypes.c_int(tr), arr, ctypes.c_int(x.size), x.ctypes.data_as(ctypes.c_void_p),
                             y.ctypes.data_as(ctypes.c_void_p))
    if rc:
        raise RefError(rc, lib().ref_last_error().decode())
    return y


def chebd1_mult(x):
    """The reference's 1-D ChebD1Mult (chebyshev.c:41-77)."""
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    y = np.zeros_like(x)
    rc = lib().ref_chebd1_mult(ctypes.c_int(x.size), x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p))
    if rc:
        raise RefError(rc, lib().ref_last_error().decode())
    return y


# ---- elliptic.C ----------------------------------------------------------------------------------------------------------
_EPATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libellipticref.so")
_elib = None


def elliptic_available():
    return os.path.exists(_EPATH)


def elib():
    global _elib
    if _elib is None:
        L = ctypes.CDLL(_EPATH)
        L.ref_elliptic_create.restype = ctypes.c_void_p
        _elib = L
    return _elib


class RefElliptic:
    """The reference's own MatCreate_Elliptic / SetupBC / CreateExactSolution / FormFunction / MatMult_Elliptic / FormJacobian
    (elliptic.C, compiled unmodified against the PETSc / FFTW stand-ins), driven the way elliptic.C's main() drives them."""

    def __init__(self, dim, gamma=0.0, exponent=2.0, exact=2, cos_scale=1.0):
        dim = [int(v) for v in dim]
        arr = (ctypes.c_int * len(dim))(*dim)
        self._h = ctypes.c_void_p(elib().ref_elliptic_create(len(dim), arr, ctypes.c_double(gamma), ctypes.c_double(exponent),
                                                             ctypes.c_int(exact), ctypes.c_double(cos_scale)))
        if not self._h:
            raise RefError(1, "MatCreate_Elliptic / CreateExactSolution failed")
        m, g, nd = ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_longlong()
        elib().ref_elliptic_sizes(self._h, ctypes.byref(m), ctypes.byref(g), ctypes.byref(nd))
        self.m, self.g, self.nd, self.d, self.dim = m.value, g.value, nd.value, len(dim), dim

    def _get(self, which, n):
        out = np.empty(n)
        if elib().ref_elliptic_get(self._h, ctypes.c_int(which), out.ctypes.data_as(ctypes.c_void_p)):
            raise RefError(1, "bad selector")
        return out

    u = property(lambda s: s._get(0, s.g))
    u2 = property(lambda s: s._get(1, s.g))
    dirichlet = property(lambda s: s._get(2, s.nd))
    b = property(lambda s: s._get(3, s.g))
    eta = property(lambda s: s._get(4, s.m))
    deta = property(lambda s: s._get(5, s.m))

    def gradu(self, k):
        return self._get(6 + k, self.m)

    def form_function(self, U):
        U = np.ascontiguousarray(U, dtype=np.float64).copy()
        F = np.empty_like(U)
        rc = elib().ref_elliptic_function(self._h, U.ctypes.data_as(ctypes.c_void_p), F.ctypes.data_as(ctypes.c_void_p))
        if rc:
            raise RefError(rc, "FormFunction")
        return F

    def mat_mult(self, U):
        U = np.ascontiguousarray(U, dtype=np.float64).copy()
        V = np.empty_like(U)
        rc = elib().ref_elliptic_matmult(self._h, U.ctypes.data_as(ctypes.c_void_p), V.ctypes.data_as(ctypes.c_void_p))
        if rc:
            raise RefError(rc, "MatMult_Elliptic")
        return V

    def jacobian(self):
        """FormJacobian's preconditioning matrix as scipy CSR (state = the last FormFunction)."""
        import scipy.sparse as sp

        cap = self.g * (1 + 2 * self.d)
        rows, cols = np.empty(cap, dtype=np.int32), np.empty(cap, dtype=np.int32)
        vals = np.empty(cap)
        n = elib().ref_elliptic_jacobian(self._h, ctypes.c_int(cap), rows.ctypes.data_as(ctypes.c_void_p), cols.ctypes.data_as(ctypes.c_void_p),
                                         vals.ctypes.data_as(ctypes.c_void_p))
        if n < 0 or n > cap:
            raise RefError(1, "FormJacobian")
        return sp.csr_matrix((vals[:n], (rows[:n], cols[:n])), shape=(self.g, self.g))

    def __del__(self):
        try:
            if self._h:
                elib().ref_elliptic_destroy(self._h)
                self._h = None
        except Exception:
            pass


# ---- stokes.C ------------------------------------------------------------------------------------------------------------
_SPATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libstokesref.so")
_slib = None


def stokes_available():
    return os.path.exists(_SPATH)


def slib():
    global _slib
    if _slib is None:
        L = ctypes.CDLL(_SPATH)
        L.ref_stokes_create.restype = ctypes.c_void_p
        L.ref_stokes_last_error.restype = ctypes.c_char_p
        _slib = L
    return _slib


class RefStokes:
    """The reference's own StokesCreate / StokesSetupDomain / StokesCreateExactSolution / StokesFunction / StokesMatMult{,VV,PV,VP,
    Schur} / StokesMatGetDiagonalSchur / StokesPressureReduceOrder / StokesPCSetUp0 (stokes.C + util.C, compiled unmodified against
    the PETSc / FFTW / CppAD stand-ins), -boundary 0."""

    def __init__(self, dim, rheology=0, hardness=1.0, exponent=1.0, regularization=1.0, gamma0=1.0, exact=2):
        dim = [int(v) for v in dim]
        arr = (ctypes.c_int * len(dim))(*dim)
        self._h = ctypes.c_void_p(slib().ref_stokes_create(len(dim), arr, ctypes.c_int(exact), ctypes.c_int(rheology), ctypes.c_double(hardness),
                                                           ctypes.c_double(exponent), ctypes.c_double(regularization), ctypes.c_double(gamma0)))
        if not self._h:
            raise RefError(1, "StokesCreate failed: " + slib().ref_stokes_last_error().decode())
        v = [ctypes.c_longlong() for _ in range(5)]
        slib().ref_stokes_sizes(self._h, *[ctypes.byref(x) for x in v])
        self.m, self.g, self.gp, self.gv, self.dv = [x.value for x in v]
        self.d, self.dim = len(dim), dim

    def _get(self, which, n):
        out = np.empty(n)
        if slib().ref_stokes_get(self._h, ctypes.c_int(which), out.ctypes.data_as(ctypes.c_void_p)):
            raise RefError(1, "bad selector")
        return out

    u = property(lambda s: s._get(0, s.g))
    u2 = property(lambda s: s._get(1, s.g))
    dirichlet = property(lambda s: s._get(2, s.dv))
    force = property(lambda s: s._get(3, s.g))
    eta = property(lambda s: s._get(4, s.m))
    deta = property(lambda s: s._get(5, s.m))

    def strain(self, j):
        return self._get(6 + j, self.m * self.d)

    def set_rheology(self, exponent, regularization):
        slib().ref_stokes_set_rheology(self._h, ctypes.c_double(exponent), ctypes.c_double(regularization))

    def _apply(self, fn, x, nout):
        x = np.ascontiguousarray(x, dtype=np.float64).copy()
        y = np.zeros(nout)
        rc = fn(self._h, x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p))
        if rc:
            raise RefError(rc, slib().ref_stokes_last_error().decode())
        return y

    def function(self, x):
        return self._apply(slib().ref_stokes_function, x, self.g)

    def mat_mult(self, x):
        return self._apply(slib().ref_stokes_matmult, x, self.g)

    def mat_mult_vv(self, x):
        return self._apply(slib().ref_stokes_matmult_vv, x, self.gv)

    def mat_mult_pv(self, x):
        return self._apply(slib().ref_stokes_matmult_pv, x, self.gp)

    def mat_mult_vp(self, x):
        return self._apply(slib().ref_stokes_matmult_vp, x, self.gv)

    def mat_mult_schur_identity(self, x):
        """StokesMatMultSchur with KSPSolve(KSPSchurVelocity) replaced by the identity (the stand-in's KSP)."""
        return self._apply(slib().ref_stokes_matmult_schur_identity, x, self.gp)

    def get_diagonal_schur(self):
        y = np.zeros(self.gp)
        rc = slib().ref_stokes_diag_schur(self._h, y.ctypes.data_as(ctypes.c_void_p))
        if rc:
            raise RefError(rc, "StokesMatGetDiagonalSchur")
        return y

    def pressure_reduce_order(self, pL):
        pL = np.ascontiguousarray(pL, dtype=np.float64).copy()
        rc = slib().ref_stokes_reduce_order(self._h, pL.ctypes.data_as(ctypes.c_void_p))
        if rc:
            raise RefError(rc, slib().ref_stokes_last_error().decode())
        return pL

    def pc_velocity_matrix(self):
        import scipy.sparse as sp

        cap = self.gv * (1 + 2 * self.d) + 16
        rows, cols = np.empty(cap, dtype=np.int32), np.empty(cap, dtype=np.int32)
        vals = np.empty(cap)
        n = slib().ref_stokes_pc_matrix(self._h, ctypes.c_int(cap), rows.ctypes.data_as(ctypes.c_void_p), cols.ctypes.data_as(ctypes.c_void_p),
                                        vals.ctypes.data_as(ctypes.c_void_p))
        if n < 0 or n > cap:
            raise RefError(1, "StokesPCSetUp0")
        return sp.csr_matrix((vals[:n], (rows[:n], cols[:n])), shape=(self.gv, self.gv))

    def __del__(self):
        try:
            if self._h:
                slib().ref_stokes_destroy(self._h)
                self._h = None
        except Exception:
            pass
