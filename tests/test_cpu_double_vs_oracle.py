"""The CPU test double of the device-side entry points (tests/mock/sb200_cpu_double.cpp) against the oracle, so that the
host-layer tests which run over the double (tests/test_native_cpu_double.py) rest on verified arithmetic.  The double is test
infrastructure only - it is never part of libspectral_b200.so."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle.elliptic import MatElliptic
from oracle.stokes import StokesCtx
from conftest import rel_max

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope="module")
def dbl(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("double") / "libsb200_cpu_double.so")
    src = ["tests/mock/sb200_cpu_double.cpp", "spectral_petsc_b200/csrc/cheb_matrix.cpp"]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out] + [os.path.join(ROOT, s) for s in src])
    return ctypes.CDLL(out)


@pytest.mark.parametrize("dim,gamma", [([8, 6], 4.0), ([7, 6, 5], 4.0), ([6, 5, 4, 5], 1.5), ([16, 16, 16], 0.0)], ids=str)
def test_elliptic_double(dbl, dim, gamma):
    O = MatElliptic(dim, gamma=gamma, exponent=2.0)
    O.create_exact_solution(2)
    h = ctypes.c_void_p()
    assert dbl.sb200_elliptic_create(len(dim), (ctypes.c_int * len(dim))(*dim), ctypes.byref(h)) == 0
    dbl.sb200_elliptic_set_params(h, ctypes.c_double(gamma), ctypes.c_double(2.0))
    dbl.sb200_elliptic_set_dirichlet(h, P(O.dirichlet), None)
    dbl.sb200_elliptic_set_rhs(h, P(O.b), None)
    rng = np.random.default_rng(0)
    Us, U = 0.1 * rng.standard_normal(O.g), rng.standard_normal(O.g)
    F, V = np.empty(O.g), np.empty(O.g)
    assert dbl.sb200_elliptic_function(h, P(Us), P(F), None) == 0
    assert rel_max(F, O.form_function(Us)) < 1e-11
    assert dbl.sb200_elliptic_matmult(h, P(U), P(V), None) == 0
    assert rel_max(V, O.mat_mult(U)) < 1e-11
    nrows, nnz = ctypes.c_longlong(), ctypes.c_longlong()
    dbl.sb200_elliptic_jacobian_sizes(h, ctypes.byref(nrows), ctypes.byref(nnz))
    rp, ci, va = np.empty(nrows.value + 1, dtype=np.int32), np.empty(nnz.value, dtype=np.int32), np.empty(nnz.value)
    assert dbl.sb200_elliptic_jacobian_csr(h, P(rp), P(ci), P(va), None) == 0
    ref = O.form_jacobian_matrix().tocsr()
    ref.sort_indices()
    assert np.array_equal(rp, ref.indptr) and np.array_equal(ci, ref.indices) and rel_max(va, ref.data) < 1e-10
    dbl.sb200_elliptic_destroy(h)


@pytest.mark.parametrize("dim,rheology", [([8, 6], 0), ([8, 6], 1), ([7, 6, 5], 1), ([10, 10, 10], 1)], ids=str)
def test_stokes_double(dbl, dim, rheology):
    d = len(dim)
    O = StokesCtx(dim, rheology=rheology, exponent=3.0, regularization=1e-2, exact=2)
    O.create_exact_solution()
    h = ctypes.c_void_p()
    assert dbl.sb200_stokes_create(d, (ctypes.c_int * d)(*dim), ctypes.byref(h)) == 0
    v = [ctypes.c_longlong() for _ in range(5)]
    dbl.sb200_stokes_sizes(h, *[ctypes.byref(x) for x in v])
    assert [x.value for x in v] == [O.m, O.g, O.gp, O.gv, O.dv]
    dbl.sb200_stokes_set_rheology(h, rheology, ctypes.c_double(1.0), ctypes.c_double(3.0), ctypes.c_double(1e-2), ctypes.c_double(1.0))
    dbl.sb200_stokes_set_dirichlet(h, P(np.ascontiguousarray(O.dirichlet.reshape(-1))), None)
    dbl.sb200_stokes_set_force(h, P(O.force), None)
    rng = np.random.default_rng(1)
    xs, x = 0.3 * rng.standard_normal(O.g), rng.standard_normal(O.g)
    F, y = np.empty(O.g), np.empty(O.g)
    assert dbl.sb200_stokes_function(h, P(xs), P(F), None) == 0
    assert rel_max(F, O.function(xs)) < 1e-10
    mn, mx = ctypes.c_double(), ctypes.c_double()
    dbl.sb200_stokes_eta_minmax(h, ctypes.byref(mn), ctypes.byref(mx), None)
    assert mn.value == pytest.approx(O.min_eta, rel=1e-12) and mx.value == pytest.approx(O.max_eta, rel=1e-12)
    assert dbl.sb200_stokes_matmult(h, P(x), P(y), None) == 0
    assert rel_max(y, O.mat_mult(x)) < 1e-10
    xv, xp = O.split(x)
    yv, yp = np.empty(O.gv), np.empty(O.gp)
    dbl.sb200_stokes_matmult_vv(h, P(xv), P(yv), None)
    assert rel_max(yv, O.mat_mult_vv(xv)) < 1e-10
    dbl.sb200_stokes_matmult_pv(h, P(xv), P(yp), None)
    assert rel_max(yp, O.mat_mult_pv(xv)) < 1e-10
    dbl.sb200_stokes_matmult_vp(h, P(xp), P(yv), None)
    assert rel_max(yv, O.mat_mult_vp(xp)) < 1e-10
    dbl.sb200_stokes_get_diagonal_schur(h, P(yp), None)
    assert rel_max(yp, O.get_diagonal_schur()) < 1e-13
    pL = rng.standard_normal(O.m)
    ref = O.pressure_reduce_order(pL.copy())
    dbl.sb200_stokes_pressure_reduce_order(h, P(pL), None)
    assert rel_max(pL, ref) < 1e-9  # high-degree extrapolation: conditioning, not a different formula
    nrows, nnz = ctypes.c_longlong(), ctypes.c_longlong()
    dbl.sb200_stokes_pc_velocity_sizes(h, ctypes.byref(nrows), ctypes.byref(nnz))
    rp, ci, va = np.empty(nrows.value + 1, dtype=np.int32), np.empty(nnz.value, dtype=np.int32), np.empty(nnz.value)
    dbl.sb200_stokes_pc_velocity_csr(h, P(rp), P(ci), P(va), None)
    refm = O.pc_velocity_matrix().tocsr()
    refm.sort_indices()
    assert np.array_equal(rp, refm.indptr) and np.array_equal(ci, refm.indices) and rel_max(va, refm.data) < 1e-10
    dbl.sb200_stokes_destroy(h)
