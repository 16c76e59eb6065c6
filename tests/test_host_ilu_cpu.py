"""sb200_host_ilu_* (host stand-in for PETSc's PCILU, elliptic.C:183-184): the level-of-fill ILU(k) against a dense
restatement of the definition, against exact LU (enough levels) and on the finite-difference matrices it is used for."""
import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla

import spectral_petsc_b200 as sp
from oracle.elliptic import MatElliptic


def dense_iluk(A, K):
    """Textbook ILU(K) (Saad, Iterative Methods for Sparse Linear Systems, Alg. 10.5): every entry of row i is updated by every
    admissible pivot, its level is lev(i,j) = min(lev(i,j), lev(i,k) + lev(k,j) + 1), and entries whose level ends above K are dropped."""
    n = A.shape[0]
    A = A.toarray()
    INF = 10 ** 9
    lev = np.where(A != 0, 0, INF)
    np.fill_diagonal(lev, 0)
    F = A.copy()
    for i in range(n):
        for k in range(i):
            if lev[i, k] > K:
                continue
            F[i, k] /= F[k, k]
            for j in range(k + 1, n):
                if lev[k, j] > K:  # dropped from row k
                    continue
                F[i, j] -= F[i, k] * F[k, j]
                lev[i, j] = min(lev[i, j], lev[i, k] + lev[k, j] + 1)
        F[i, lev[i] > K] = 0.0
    return F, lev <= K


def to_dense(ilu):
    rp, ci, v = ilu.factor()
    return sps.csr_matrix((v, ci, rp), shape=(ilu.n, ilu.n)).toarray(), sps.csr_matrix((np.ones_like(v), ci, rp), shape=(ilu.n, ilu.n)).toarray() > 0


def fd_matrix(dim, gamma=4.0):
    O = MatElliptic(dim, gamma=gamma, exponent=2.0)
    O.form_function(0.1 * np.random.default_rng(1).standard_normal(O.g))
    P = O.form_jacobian_matrix().tocsr()
    P.sort_indices()
    return P


@pytest.mark.parametrize("K", [0, 1, 2, 3])
def test_factor_equals_the_definition(K):
    rng = np.random.default_rng(K)
    A = sps.random(40, 40, density=0.08, random_state=7, format="csr") + sps.identity(40) * 4.0
    A = A.tocsr()
    A.sort_indices()
    for M in (A, fd_matrix([6, 5, 4]), fd_matrix([7, 6])):
        ilu = sp.HostILU(M, K)
        F, pat = to_dense(ilu)
        Fd, patd = dense_iluk(M, K)
        assert np.array_equal(pat, patd)
        assert np.abs(F - Fd).max() <= 1e-13 * np.abs(Fd).max()
        b = rng.standard_normal(M.shape[0])
        L = np.tril(Fd, -1) + np.eye(M.shape[0])
        U = np.triu(Fd)
        assert np.allclose(ilu.solve(b), np.linalg.solve(U, np.linalg.solve(L, b)), rtol=1e-11, atol=1e-12)


def test_level_zero_keeps_the_pattern_and_many_levels_give_lu():
    P = fd_matrix([6, 6, 6])
    ilu0 = sp.HostILU(P, 0)
    rp, ci, _ = ilu0.factor()
    assert np.array_equal(rp, P.indptr) and np.array_equal(ci, P.indices) and ilu0.nnz == P.nnz
    nnz = [sp.HostILU(P, k).nnz for k in range(4)]
    assert nnz == sorted(nnz) and nnz[2] > nnz[0]
    full = sp.HostILU(P, 10 ** 6)  # every fill entry allowed: the exact LU
    b = np.random.default_rng(0).standard_normal(P.shape[0])
    assert np.allclose(full.solve(b), spla.spsolve(P.tocsc(), b), rtol=1e-10, atol=1e-12)
    # ILU(2) is a better preconditioner than ILU(0): ||I - M^-1 P|| on random vectors
    r = lambda ilu: np.linalg.norm(b - ilu.solve(P @ b)) / np.linalg.norm(b)
    assert r(sp.HostILU(P, 2)) < r(ilu0) < 1.0


def test_refactor_and_errors():
    P = fd_matrix([6, 5, 4])
    ilu = sp.HostILU(P, 2)
    Q = P.copy()
    Q.data = Q.data * 1.5
    ilu.refactor(Q)
    F, _ = to_dense(ilu)
    Fd, _ = dense_iluk(Q, 2)
    assert np.abs(F - Fd).max() <= 1e-13 * np.abs(Fd).max()
    Z = sps.csr_matrix(np.array([[0.0, 1.0], [1.0, 0.0]]))
    with pytest.raises(sp.SB200Error) as ei:
        sp.HostILU(Z, 0)
    assert ei.value.code == 83 and "zero pivot" in str(ei.value)
