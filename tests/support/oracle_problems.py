"""Problem adapters over the CPU oracle with the interface of spectral_petsc_b200.drivers.GpuElliptic / GpuStokes, so the
driver flows (option handling, Newton / continuation loops, printed lines, VTK) can be run without a GPU and compared with
the same flows on the CUDA shells.  Test infrastructure only."""
import numpy as np

from oracle.elliptic import MatElliptic
from oracle.fgmres import fgmres
from oracle.stokes import StokesCtx


def np_krylov(op, b, pc, rtol, maxits, restart):
    x, its, hist, reason = fgmres(op, b, M=pc, restart=restart, rtol=rtol, maxits=maxits)
    return x, its, reason


class OracleElliptic:
    def __init__(self, dim, gamma, exponent):
        self.O = O = MatElliptic(dim, gamma=gamma, exponent=exponent)
        self.m, self.g, self.nd = O.m, O.g, O.nd
        self.krylov = np_krylov
        self.form_function, self.mat_mult, self.jacobian = O.form_function, O.mat_mult, O.form_jacobian_matrix

    def from_host(self, a):
        return np.array(a, dtype=np.float64, copy=True)

    def to_host(self, v):
        return v

    def set_dirichlet(self, a):
        self.O.dirichlet = np.array(a, copy=True)

    def set_rhs(self, a):
        self.O.b = np.array(a, copy=True)


class OracleStokes:
    def __init__(self, dim, rheology, hardness, exponent, regularization, gamma0):
        self.O = O = StokesCtx(dim, rheology=rheology, hardness=hardness, exponent=exponent, regularization=regularization, gamma0=gamma0)
        self.d, self.dim = O.d, list(dim)
        self.m, self.g, self.gp, self.gv, self.dv = O.m, O.g, O.gp, O.gv, O.dv
        self.krylov = np_krylov
        self.mat_mult, self.mat_mult_vv, self.mat_mult_pv, self.mat_mult_vp = O.mat_mult, O.mat_mult_vv, O.mat_mult_pv, O.mat_mult_vp
        self.get_diagonal_schur, self.function, self.set_rheology = O.get_diagonal_schur, O.function, O.set_rheology
        self.pc_velocity_matrix = O.pc_velocity_matrix

    def from_host(self, a):
        return np.array(a, dtype=np.float64, copy=True)

    def to_host(self, v):
        return v

    def set_dirichlet(self, a):
        self.O.dirichlet = np.array(a, copy=True).reshape(-1, self.d)

    def set_force(self, a):
        self.O.force = np.array(a, copy=True)

    def eta_minmax(self):
        return self.O.min_eta, self.O.max_eta

    def state_host(self):
        return self.O.eta, self.O.deta, self.O.strain

    def pressure_reduce_order_host(self, pL):
        return self.O.pressure_reduce_order(pL)
