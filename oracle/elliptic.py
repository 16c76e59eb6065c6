"""Oracle restatement of the elliptic.C operator path (test infrastructure).

Follows /root/reference/elliptic.C:
  MatCreate_Elliptic  :250-293  contexts, coordinates, work vectors
  SetupBC             :372-466  lexicographic walk, interior -> global id, boundary -> dirichlet id
  MatMult_Elliptic    :297-339  Jacobian action
  FormFunction        :481-533  nonlinear residual (caches gradu, eta, deta)
  FormJacobian        :537-590  finite-difference preconditioning matrix (PC input; not on the hot path)
  CreateExactSolution :594-677  -exact 0/1/2 manufactured solutions
"""
import math

import numpy as np

from .chebyshev import ChebCtx, PI


class EllipticError(ValueError):
    pass


class MatElliptic:
    """The MatElliptic context (elliptic.C:78-86) with the reference's state."""

    def __init__(self, dim, gamma=0.0, exponent=2.0, workers=1):
        self.d = d = len(dim)
        self.dim = [int(x) for x in dim]
        self.m = m = int(np.prod(self.dim))
        self.gamma, self.exponent = float(gamma), float(exponent)
        # :270 one Chebyshev derivative context per axis
        self.D = [ChebCtx(d, i, self.dim, m, workers=workers) for i in range(d)]
        self.eta = np.ones(m)  # :266
        self.deta = np.zeros(m)  # :267
        self.gradu = [np.zeros(m) for _ in range(d)]
        # :275-281 coordinates, block size d
        idx = np.indices(self.dim).reshape(d, -1)
        self.x = np.empty((m, d))
        for j in range(d):
            self.x[:, j] = np.cos(idx[j] * math.pi / (self.dim[j] - 1))
        # SetupBC :386-415. Walk order = row-major flat index. Boundary iff any index on an end.
        on_bdy = np.zeros(m, dtype=bool)
        for j in range(d):
            on_bdy |= (idx[j] == 0) | (idx[j] == self.dim[j] - 1)
        self.ixG = np.flatnonzero(~on_bdy)  # global id -> local index (:409)
        self.ixD = np.flatnonzero(on_bdy)  # dirichlet id -> local index (:404)
        self.ixL = np.full(m, -1, dtype=np.int64)  # local -> global id or -1 (:403,408)
        self.ixL[self.ixG] = np.arange(self.ixG.size)
        self.g = self.ixG.size
        self.nd = self.ixD.size
        self.dirichlet = np.zeros(self.nd)  # DirichletBdy gives value 0 (:470-477)
        self.dirichlet0 = np.zeros(self.nd)  # :463-464
        self.b = np.zeros(self.g)

    # -- the hot path ---------------------------------------------------
    def mat_mult(self, U):
        """MatMult_Elliptic (elliptic.C:297-339)."""
        d, m = self.d, self.m
        w0 = np.zeros(m)
        w0[self.ixG] = U  # scatterGL :305
        w0[self.ixD] = self.dirichlet0  # scatterDL :307
        w = [self.D[k].mult(w0) for k in range(d)]  # :309-311
        for k in range(d):  # :319-323
            w[k] = self.eta * w[k] + self.deta * w0 * self.gradu[k]
        out = np.zeros(m)  # :329
        for k in range(d):  # :330-333 accumulate in axis order
            out = out + (-1.0) * self.D[k].mult(w[k])
        return out[self.ixG]  # scatterLG :336

    def form_function(self, U):
        """FormFunction (elliptic.C:481-533).  Updates gradu/eta/deta caches."""
        d, m = self.d, self.m
        w0 = np.zeros(m)
        w0[self.ixG] = U  # :489
        w0[self.ixD] = self.dirichlet  # :491
        for k in range(d):  # :497-499
            self.gradu[k] = self.D[k].mult(w0)
        with np.errstate(invalid="ignore", divide="ignore"):
            self.eta = 1.0 + self.gamma * np.power(w0, self.exponent)  # :508
            self.deta = self.exponent * self.gamma * np.power(w0, self.exponent - 1.0)  # :509
        w = [self.eta * self.gradu[k] for k in range(d)]  # :511
        out = np.zeros(m)
        for k in range(d):  # :521-524
            out = out + (-1.0) * self.D[k].mult(w[k])
        return out[self.ixG] + (-1.0) * self.b  # :529-531

    # -- fixtures -------------------------------------------------------
    def create_exact_solution(self, exact, cos_scale=None):
        """CreateExactSolution (elliptic.C:594-677): returns (u, u2), sets dirichlet and b."""
        d, m = self.d, self.m
        X = self.x
        gamma, exponent = self.gamma, self.exponent
        s = 0.5
        if exact in (0, 3):
            if cos_scale is None:
                raise EllipticError("-cos_scale has no default in the reference (elliptic.C:607-609)")
            s *= cos_scale
        w0 = np.empty(m)
        w1 = np.empty(m)
        if exact == 0:  # :620-632
            v = np.ones(m)
            for j in range(d):
                v = v * np.cos(s * PI * X[:, j])
            with np.errstate(invalid="ignore", divide="ignore"):
                eta = 1.0 + gamma * np.power(v, exponent)
                deta = np.zeros(m) if abs(exponent) < 1e-10 else gamma * exponent * np.power(v, exponent - 1.0)
            acc = np.zeros(m)
            for j in range(d):
                dv = np.ones(m)
                for k in range(d):
                    dv = dv * (-s * PI * np.sin(s * PI * X[:, k]) if k == j else np.cos(s * PI * X[:, k]))
                d2v = -((s * PI) ** 2) * v
                acc = acc + (deta * dv ** 2 + eta * d2v)
            w0[:] = v
            w1[:] = -acc
        elif exact == 1:  # :633-643
            v = np.ones(m)
            acc = np.zeros(m)
            for j in range(d):
                v = v * ((1 - X[:, j]) * (1 + X[:, j]))
                z = np.ones(m)
                for k in range(d):
                    if k != j:
                        z = z * (2.0 * (1 - X[:, k]) * (1 + X[:, k]))
                acc = acc + z
            w0[:] = v
            w1[:] = acc
        elif exact == 2:  # :644-655
            v = np.ones(m)
            acc = np.zeros(m)
            for j in range(d):
                v = v * np.power(X[:, j], 4 + j)
                z = np.ones(m)
                for k in range(d):
                    if k == j:
                        z = z * ((4 + k) * (3 + k) * np.power(X[:, k], 2 + k))
                    else:
                        z = z * np.power(X[:, k], 4 + k)
                acc = acc - z
            w0[:] = v
            w1[:] = acc
        else:
            raise EllipticError("Choose an exact solution.")  # :657
        u = w0[self.ixG].copy()  # :668
        u2 = w1[self.ixG].copy()  # :670
        self.dirichlet = w0[self.ixD].copy()  # :672
        self.b = u2.copy()  # :674
        return u, u2

    def form_jacobian_matrix(self):
        """FormJacobian (elliptic.C:537-590): the FD preconditioning matrix P as scipy CSR.

        Out of the hot path (it feeds PETSc's PC); restated so solver-level tests can
        use the same PC input as the reference would.
        """
        import scipy.sparse as sp

        d, m, dim = self.d, self.m, self.dim
        strides = [int(np.prod(dim[j + 1:])) for j in range(d)]
        rows, cols, vals = [], [], []
        x, eta, deta, ixL = self.x, self.eta, self.deta, self.ixL
        I = self.ixG  # local indices of interior nodes, in global order
        diag = np.zeros(I.size)
        for j in range(d):
            iM = I - strides[j]
            iP = I + strides[j]
            x0, xMM, xPP = x[I, j], x[iM, j], x[iP, j]
            xM = 0.5 * (xMM + x0)
            idxM = 1.0 / (x0 - xMM)
            xP = 0.5 * (x0 + xPP)
            idxP = 1.0 / (xPP - x0)
            idx = 1.0 / (xP - xM)
            eM = 0.5 * (eta[iM] + eta[I]); deM = 0.5 * (deta[iM] + deta[I]); du0M = 0.5 * (self.gradu[j][iM] + self.gradu[j][I])
            eP = 0.5 * (eta[iP] + eta[I]); deP = 0.5 * (deta[iP] + deta[I]); du0P = 0.5 * (self.gradu[j][iP] + self.gradu[j][I])
            vM = -idx * (idxM * eM - 0.5 * deM * du0M)
            vP = -idx * (idxP * eP + 0.5 * deP * du0P)
            diag = diag + idx * (idxP * eP + idxM * eM - 0.5 * (deP * du0P - deM * du0M))
            for nb, v in ((iM, vM), (iP, vP)):
                gnb = ixL[nb]
                keep = gnb >= 0  # MatSetValues ignores negative column indices
                rows.append(np.arange(I.size)[keep]); cols.append(gnb[keep]); vals.append(v[keep])
        rows.append(np.arange(I.size)); cols.append(np.arange(I.size)); vals.append(diag)
        return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(I.size, I.size))
