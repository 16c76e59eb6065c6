"""Oracle restatement of the stokes.C operator path, Dirichlet case (test infrastructure).

Follows /root/reference/stokes.C and util.C:
  StokesCreate / StokesSetupDomain (Dirichlet branch)  :257-345, :773-938
  StokesMatMult            :499-519     StokesMatMultSchur / GetDiagonalSchur  :523-553
  StokesDivergence / PV    :557-595     StokesMatMultVP                        :599-619
  StokesMatMultVV          :623-676     StokesFunction                         :680-758
  StokesPressureReduceOrder:1029-1080   polyInterp (util.C)                    :129-144
  StokesRheologyLinear/Power :1920-1944 continuation parameters                :214-221
  StokesExact0..3 :1948-2034            StokesCreateExactSolution              :942-1003
  StokesPCSetUp0 (FD velocity matrix, PC input, off the hot path)              :1160-1240
  StokesPCApply0 block-LU orchestration                                        :1714-1745

Layouts: local velocity is AoS [i0][i1][i2][k] (index i*d+k, :651), local pressure [i0][i1][i2];
the global vector is AoS per interior node [v_0..v_{d-1}, p] (:867-877); dirichlet is boundary
nodes in walk order times d components (:796-801).
"""
import math

import numpy as np

from .chebyshev import ChebCtx


class StokesError(ValueError):
    pass


def poly_interp(n, x, f, x0, x1):
    """polyInterp (util.C:129-144): Neville tableau evaluated at x0 and x1.

    x: (n,) nodes; f: (n, L) values for L lines sharing the nodes.  Returns (f0, f1) of shape (L,).
    The reference ping-pongs between column pairs of a width-4 work array; the arithmetic per
    column step is restated exactly (same operand order).
    """
    T0 = np.array(f, dtype=np.float64, copy=True)
    T1 = T0.copy()
    for di in range(1, n):
        cnt = n - di
        xi = x[:cnt].reshape(-1, *([1] * (T0.ndim - 1)))
        xid = x[di:di + cnt].reshape(-1, *([1] * (T0.ndim - 1)))
        den = xi - xid
        T0n = ((x0 - xid) * T0[:cnt] + (xi - x0) * T0[1:cnt + 1]) / den
        T1n = ((x1 - xid) * T1[:cnt] + (xi - x1) * T1[1:cnt + 1]) / den
        T0[:cnt] = T0n
        T1[:cnt] = T1n
    return T0[0], T1[0]


def rheology_linear(gamma):
    return np.ones_like(gamma), np.zeros_like(gamma)  # :1920-1926


def rheology_power(gamma, hardness, exponent, regularization, gamma0):
    """StokesRheologyPower (stokes.C:1930-1944)."""
    n = exponent
    p = (1.0 - n) / (2.0 * n)
    base = regularization + gamma / gamma0
    eta = hardness * np.power(base, p)
    if abs(n) > 1.0e-5:
        deta = hardness * p / gamma0 * np.power(base, p - 1.0)
    else:
        deta = np.zeros_like(gamma)
    return eta, deta


def continuation_params(i, cont, exponent, regularization):
    """stokes.C:218-219."""
    e = 1.0 + math.pow(1.0 * i / cont, 0.8) * (exponent - 1.0)
    r = math.exp(math.log(regularization) * i / cont)
    return e, r


class StokesCtx:
    def __init__(self, dim, rheology=0, hardness=1.0, exponent=1.0, regularization=1.0, gamma0=1.0, exact=0, workers=1):
        self.d = d = len(dim)
        if d not in (2, 3):
            raise StokesError("oracle restates the d = 2, 3 paths (StokesPressureReduceOrder, stokes.C:1036)")
        self.dim = [int(x) for x in dim]
        self.m = m = int(np.prod(self.dim))
        self.rheology = rheology
        self.hardness, self.exponent, self.regularization, self.gamma0 = hardness, exponent, regularization, gamma0
        self.exact = exact
        # :284-291 scalar (DP) and vector (DV) derivative contexts
        self.DP = [ChebCtx(d, i, self.dim, m, workers=workers) for i in range(d)]
        self.DV = [ChebCtx(d + 1, i, self.dim + [d], m * d, workers=workers) for i in range(d)]
        idx = np.indices(self.dim).reshape(d, -1)
        self.coord = np.empty((m, d))
        for j in range(d):
            self.coord[:, j] = np.cos(idx[j] * math.pi / (self.dim[j] - 1))  # :296
        on_bdy = np.zeros(m, dtype=bool)
        for j in range(d):
            on_bdy |= (idx[j] == 0) | (idx[j] == self.dim[j] - 1)
        self.int_nodes = np.flatnonzero(~on_bdy)  # local node index of interior nodes, walk order
        self.bdy_nodes = np.flatnonzero(on_bdy)
        self.gp = self.int_nodes.size
        self.gv = self.gp * d
        self.g = self.gp * (d + 1)
        self.dv = self.bdy_nodes.size * d
        self.eta = np.ones(m)
        self.deta = np.zeros(m)
        self.strain = [np.zeros((m, d)) for _ in range(d)]
        self.force = np.zeros(self.g)
        # StokesDirichlet evaluates the exact solution at boundary nodes (:2039-2050, :796)
        self.dirichlet = np.zeros((self.bdy_nodes.size, d))
        for q, node in enumerate(self.bdy_nodes):
            val, _ = self.exact_fn(self.coord[node])
            self.dirichlet[q] = val[:d]

    # ---- exact solutions ------------------------------------------------------
    def exact_fn(self, c):
        """StokesExact0..3 (stokes.C:1948-2034): returns (value[d+1], rhs[d+1]).

        For -exact 2 in 3-D the reference leaves value[3] (pressure) unset; the oracle
        defines it as 0.
        """
        d = self.d
        val = np.zeros(d + 1)
        rhs = np.zeros(d + 1)
        if self.exact == 0:
            return val, rhs
        if self.exact in (1, 2):
            eta = 1.0
            u = math.sin(0.5 * math.pi * c[0]) * math.cos(0.5 * math.pi * c[1])
            v = -math.cos(0.5 * math.pi * c[0]) * math.sin(0.5 * math.pi * c[1])
            val[0], val[1] = u, v
            rhs[0] = (0.5 * math.pi) ** 2 * eta * u
            rhs[1] = (0.5 * math.pi) ** 2 * eta * v
            if self.exact == 1:
                val[d] = 0.25 * (math.cos(math.pi * c[0]) + math.cos(math.pi * c[1])) + 10 * (c[0] + c[1])
                rhs[0] += -0.25 * math.pi * math.sin(math.pi * c[0]) + 10
                rhs[1] += -0.25 * math.pi * math.sin(math.pi * c[1]) + 10
            return val, rhs
        if self.exact == 3:
            if d != 2:
                raise StokesError("StokesExact3 only implemented for dimension 2")
            val[0] = c[1] + 1.0
            return val, rhs
        raise StokesError("Exact solution %d not implemented" % self.exact)

    def create_exact_solution(self):
        """StokesCreateExactSolution (stokes.C:942-1003): returns (U, U2) and sets force = U2."""
        d = self.d
        U = np.zeros((self.gp, d + 1))
        U2 = np.zeros((self.gp, d + 1))
        for q, node in enumerate(self.int_nodes):
            val, rhs = self.exact_fn(self.coord[node])
            U[q] = val
            U2[q] = rhs
        self.force = U2.reshape(-1).copy()
        return U.reshape(-1), U2.reshape(-1)

    # ---- scatters -------------------------------------------------------------
    def split(self, xG):
        X = xG.reshape(self.gp, self.d + 1)
        return X[:, :self.d].reshape(-1).copy(), X[:, self.d].copy()  # scatterGV, scatterGP

    def merge(self, vG, pG):
        X = np.empty((self.gp, self.d + 1))
        X[:, :self.d] = vG.reshape(self.gp, self.d)
        X[:, self.d] = pG
        return X.reshape(-1)

    def vel_local(self, vG, with_dirichlet):
        xL = np.zeros((self.m, self.d))
        xL[self.int_nodes] = vG.reshape(self.gp, self.d)  # scatterVL
        if with_dirichlet:
            xL[self.bdy_nodes] = self.dirichlet  # scatterDL
        return xL

    def dvel(self, axis, xL):
        return self.DV[axis].mult(xL.reshape(-1)).reshape(self.m, self.d)

    def dpres(self, axis, pL):
        return self.DP[axis].mult(pL.reshape(-1))

    # ---- shells ---------------------------------------------------------------
    def mat_mult_vv(self, xG):
        """StokesMatMultVV (stokes.C:623-676)."""
        d = self.d
        xL = self.vel_local(xG, False)
        V = [self.dvel(i, xL) for i in range(d)]  # V[j][:, k] = d_j u_k
        z = np.zeros(self.m)
        strain = [[None] * d for _ in range(d)]
        for j in range(d):
            for k in range(d):
                strain[j][k] = 0.5 * (V[j][:, k] + V[k][:, j])  # :651
                z = z + strain[j][k] * self.strain[j][:, k]  # :652
        for j in range(d):
            Vj = np.empty((self.m, d))
            for k in range(d):
                s = self.eta * strain[j][k]  # :657
                Vj[:, k] = s + self.deta * self.strain[j][:, k] * z  # :659
            V[j] = Vj
        yL = np.zeros((self.m, d))
        for i in range(d):  # :668-671
            yL = yL + (-1.0) * self.dvel(i, V[i])
        return yL[self.int_nodes].reshape(-1)

    def divergence(self, with_dirichlet, xG):
        """StokesDivergence (stokes.C:570-595)."""
        xL = self.vel_local(xG, with_dirichlet)
        acc = np.zeros(self.m)
        for i in range(self.d):
            acc = acc + self.dpres(i, np.ascontiguousarray(xL[:, i]))
        return acc[self.int_nodes]

    def mat_mult_pv(self, xG):
        return self.divergence(False, xG)  # :557-566

    def pressure_reduce_order(self, pres):
        """StokesPressureReduceOrder (stokes.C:1029-1080) on a local pressure array (modified in place)."""
        d, dim = self.d, self.dim
        m, n = dim[0], dim[1]
        p = 1 if d == 2 else dim[2]
        Pz = pres.reshape(m, n, p)
        C = self.coord.reshape(m, n, p, d)
        # The i-loop interleaves the z and y passes plane by plane; planes are independent, so
        # the passes are applied to all i in 1..m-1 at once (i = m-1 is later overwritten by the x pass).
        if p > 1:
            xs = C[0, 0, 1:p - 1, 2]  # z nodes (same for every line)
            f = np.moveaxis(Pz[1:m, 1:n, 1:p - 1], 2, 0)  # (p-2, m-1, n-1)
            f0, f1 = poly_interp(p - 2, xs, f, C[0, 0, 0, 2], C[0, 0, p - 1, 2])
            Pz[1:m, 1:n, 0] = f0
            Pz[1:m, 1:n, p - 1] = f1
        xs = C[0, 1:n - 1, 0, 1]
        f = np.moveaxis(Pz[1:m, 1:n - 1, :], 1, 0)  # (n-2, m-1, p)
        f0, f1 = poly_interp(n - 2, xs, f, C[0, 0, 0, 1], C[0, n - 1, 0, 1])
        Pz[1:m, 0, :] = f0
        Pz[1:m, n - 1, :] = f1
        xs = C[1:m - 1, 0, 0, 0]
        f = Pz[1:m - 1, :, :]  # (m-2, n, p)
        f0, f1 = poly_interp(m - 2, xs, f, C[0, 0, 0, 0], C[m - 1, 0, 0, 0])
        Pz[0, :, :] = f0
        Pz[m - 1, :, :] = f1
        return pres

    def mat_mult_vp(self, pG):
        """StokesMatMultVP (stokes.C:599-619)."""
        pL = np.zeros(self.m)
        pL[self.int_nodes] = pG
        self.pressure_reduce_order(pL)
        vL = np.zeros((self.m, self.d))
        for i in range(self.d):
            vL[:, i] = self.dpres(i, pL)
        return vL[self.int_nodes].reshape(-1)

    def mat_mult(self, xG):
        """StokesMatMult (stokes.C:499-519)."""
        v, p = self.split(xG)
        vG1 = self.mat_mult_vv(v)
        pG1 = self.mat_mult_pv(v)
        vG0 = self.mat_mult_vp(p)
        vG1 = vG1 + 1.0 * vG0
        return self.merge(vG1, pG1)

    def get_diagonal_schur(self):
        return 1.0 / self.eta[self.int_nodes]  # :542-553

    def mat_mult_schur(self, pG, velocity_solve):
        """StokesMatMultSchur (stokes.C:523-535); velocity_solve(rhs) stands for KSPSolve(KSPSchurVelocity)."""
        v0 = self.mat_mult_vp(pG)
        v1 = velocity_solve(v0)
        return -1.0 * self.mat_mult_pv(v1)

    def set_rheology(self, exponent, regularization):
        self.exponent, self.regularization = exponent, regularization

    def function(self, xG):
        """StokesFunction (stokes.C:680-758).  Updates strain / eta / deta caches."""
        d = self.d
        vG0, pG0 = self.split(xG)
        xL = self.vel_local(vG0, True)
        raw = [self.dvel(i, xL) for i in range(d)]  # :701
        s = [[None] * d for _ in range(d)]
        gamma = np.zeros(self.m)
        for j in range(d):
            for k in range(d):
                s[j][k] = 0.5 * (raw[j][:, k] + raw[k][:, j])  # :714
                gamma = gamma + 0.5 * (s[j][k] * s[j][k])  # :715
        if self.rheology == 0:
            self.eta, self.deta = rheology_linear(gamma)
        else:
            self.eta, self.deta = rheology_power(gamma, self.hardness, self.exponent, self.regularization, self.gamma0)
        V = []
        for j in range(d):
            Vj = np.empty((self.m, d))
            Sj = np.empty((self.m, d))
            for k in range(d):
                Vj[:, k] = self.eta * s[j][k]  # :721
                Sj[:, k] = s[j][k]  # :722
            V.append(Vj)
            self.strain[j] = Sj
        self.min_eta, self.max_eta = self.eta.min(), self.eta.max()  # :731-734
        yL = np.zeros((self.m, d))
        for i in range(d):  # :737-740
            yL = yL + (-1.0) * self.dvel(i, V[i])
        vG1 = yL[self.int_nodes].reshape(-1)
        pG1 = self.divergence(True, vG0)  # :746
        g = self.mat_mult_vp(pG0)  # :747
        vG1 = vG1 + 1.0 * g
        return self.merge(vG1, pG1) + (-1.0) * self.force  # :756

    # ---- preconditioner inputs (off the hot path) -------------------------------
    def pc_velocity_matrix(self):
        """StokesPCSetUp0 (stokes.C:1160-1240), Dirichlet case: FD velocity matrix as scipy CSR (gv x gv)."""
        import scipy.sparse as sp

        d, dim, m = self.d, self.dim, self.m
        strides = [int(np.prod(dim[j + 1:])) for j in range(d)]
        ixLnode = np.full(m, -1, dtype=np.int64)
        ixLnode[self.int_nodes] = np.arange(self.gp)
        I = self.int_nodes
        rows, cols, vals = [], [], []
        diag = np.zeros(I.size)
        x, eta = self.coord, self.eta
        nbr = []
        for j in range(d):
            iM, iP = I - strides[j], I + strides[j]
            x0, xMM, xPP = x[I, j], x[iM, j], x[iP, j]
            xM = 0.5 * (xMM + x0); idxM = 1.0 / (x0 - xMM); xP = 0.5 * (x0 + xPP); idxP = 1.0 / (xPP - x0); idx = 1.0 / (xP - xM)
            eM = 0.5 * (eta[iM] + eta[I]); eP = 0.5 * (eta[iP] + eta[I])
            nbr.append((ixLnode[iM], -idx * (idxM * eM)))
            nbr.append((ixLnode[iP], -idx * (idxP * eP)))
            diag = diag + idx * (idxP * eP + idxM * eM)
        q = np.arange(I.size)
        for f in range(d):
            rows.append(q * d + f); cols.append(q * d + f); vals.append(diag)
            for gn, v in nbr:
                keep = gn >= 0
                rows.append((q * d + f)[keep]); cols.append((gn * d + f)[keep]); vals.append(v[keep])
        return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(self.gv, self.gv))
