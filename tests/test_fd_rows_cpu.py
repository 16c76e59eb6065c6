"""The finite-difference preconditioning matrices (FormJacobian elliptic.C:537-590, StokesPCSetUp0 stokes.C:1160-1240):
the row functions the CUDA kernel runs (spectral_petsc_b200/csrc/fd_rows.h) are compiled here with g++ and driven by a
plain loop, and the CSR they produce is compared with the oracle's matrices (which tests/test_oracle_ref_*.py pin against
the reference source).  Checks the closed-form row offsets, the sorted column order and the stencil values on CPU; the
GPU kernel itself is compared in tests/test_gpu_fd_assembly.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sps

from oracle.elliptic import MatElliptic
from oracle.stokes import StokesCtx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fdlib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("fdrows") / "libfdrows_host.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", out,
                           os.path.join(ROOT, "tests", "cpp", "fd_rows_host.cpp")])
    L = ctypes.CDLL(out)
    L.fd_host_sizes.restype = ctypes.c_longlong
    return L


def host_csr(L, dim, ncomp, eta, deta, gradu, unrolled=1):
    d = len(dim)
    arr = (ctypes.c_int * d)(*dim)
    nrows = ctypes.c_longlong()
    nnz = L.fd_host_sizes(d, arr, ncomp, ctypes.byref(nrows))
    rowptr = np.full(nrows.value + 1, -1, dtype=np.int32)
    colidx = np.full(nnz, -1, dtype=np.int32)
    vals = np.full(nnz, np.nan)
    p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    g = None if gradu is None else np.ascontiguousarray(np.stack(gradu))
    rc = L.fd_host_assemble(d, arr, ncomp, p(eta), p(deta), p(g), p(rowptr), p(colidx), p(vals), unrolled)
    assert rc == 0, "row offset outside the matrix"
    return rowptr, colidx, vals


def check_csr(rowptr, colidx, vals, ref, tol=1e-14):
    ref = ref.tocsr().copy()
    ref.sort_indices()
    assert rowptr[0] == 0 and rowptr[-1] == ref.nnz == colidx.size
    assert np.array_equal(rowptr, ref.indptr)  # same entries per row, hence no gap and no overlap in the closed-form offsets
    assert np.array_equal(colidx, ref.indices)  # increasing columns within each row
    assert not np.isnan(vals).any()
    assert np.abs(vals - ref.data).max() <= tol * np.abs(ref.data).max()


ELL = [[8, 6], [3, 3], [3, 7], [7, 3, 5], [4, 4, 4], [16, 16, 16], [5, 4, 3, 6], [6] * 5, [3, 4, 3, 5, 3], [9]]


@pytest.mark.parametrize("dim", ELL, ids=str)
@pytest.mark.parametrize("gamma", [0.0, 4.0])
def test_elliptic_jacobian_rows_equal_oracle(fdlib, dim, gamma):
    O = MatElliptic(dim, gamma=gamma, exponent=2.0)
    O.form_function(0.1 * np.random.default_rng(1).standard_normal(O.g))  # eta, deta, gradu about a random state
    for unrolled in (0, 1):
        rowptr, colidx, vals = host_csr(fdlib, dim, 1, O.eta, O.deta, O.gradu, unrolled)
        check_csr(rowptr, colidx, vals, O.form_jacobian_matrix())


@pytest.mark.parametrize("dim", [[8, 6], [3, 5], [7, 6, 5], [3, 3, 3], [12, 12, 12]], ids=str)
@pytest.mark.parametrize("rheology", [0, 1])
def test_stokes_velocity_pc_rows_equal_oracle(fdlib, dim, rheology):
    S = StokesCtx(dim, rheology=rheology, exponent=3.0, regularization=1e-4, exact=2)
    S.create_exact_solution()
    S.function(np.random.default_rng(2).standard_normal(S.g))  # eta about a random state
    rowptr, colidx, vals = host_csr(fdlib, dim, S.d, S.eta, None, None)
    check_csr(rowptr, colidx, vals, S.pc_velocity_matrix())


def test_entry_count_closed_form(fdlib):
    for dim in ([128, 128, 128], [12] * 5, [20, 20, 20]):
        d = len(dim)
        nrows = ctypes.c_longlong()
        nnz = fdlib.fd_host_sizes(d, (ctypes.c_int * d)(*dim), 1, ctypes.byref(nrows))
        n = [v - 2 for v in dim]
        g = int(np.prod(n))
        assert nrows.value == g
        assert nnz == g + sum(2 * (n[j] - 1) * g // n[j] for j in range(d))
    # 128^3: 2,000,376 rows (SURVEY 8 header), 7-point stencil
    assert nrows.value == 5832 and nnz == 5832 + 3 * 2 * 17 * 18 * 18


def test_row_offsets_random_grids(fdlib):
    """Property check of the closed-form CSR offsets on random grids up to 6-D (extents 3..7, i.e. interior extents 1..5):
    rows tile [0, nnz) exactly, columns increase, every row holds its diagonal and only walk-order neighbours."""
    rng = np.random.default_rng(11)
    for _ in range(40):
        d = int(rng.integers(1, 7))
        dim = [int(v) for v in rng.integers(3, 8, size=d)]
        m = int(np.prod(dim))
        n = [v - 2 for v in dim]
        g = int(np.prod(n))
        eta = 1.0 + rng.random(m)
        for ncomp in (1, 2):
            rowptr, colidx, vals = host_csr(fdlib, dim, ncomp, eta, None, None, unrolled=int(rng.integers(0, 2)))
            assert rowptr[0] == 0 and rowptr[-1] == colidx.size and (np.diff(rowptr) >= 1).all()
            assert (colidx >= 0).all() and (colidx < g * ncomp).all() and not np.isnan(vals).any()
            istr = [int(np.prod(n[j + 1:])) for j in range(d)]
            for r in rng.integers(0, g * ncomp, size=min(50, g * ncomp)):
                cols = colidx[rowptr[r]:rowptr[r + 1]]
                assert (np.diff(cols) > 0).all() and r in cols
                node, f = divmod(int(r), ncomp)
                offs = sorted(set(abs(int(c) // ncomp - node) for c in cols if c != r))
                assert all(int(c) % ncomp == f for c in cols) and all(o in istr for o in offs)
            # symmetric positive weights: off-diagonals negative, diagonal = -(sum of ALL 2d neighbour weights) > |row sum of kept ones|
            A = sps.csr_matrix((vals, colidx, rowptr), shape=(g * ncomp, g * ncomp))
            assert (A.diagonal() > 0).all() and (A - sps.diags(A.diagonal())).max() <= 0
            assert (np.asarray(A.sum(axis=1)).ravel() >= -1e-9 * A.diagonal()).all()
