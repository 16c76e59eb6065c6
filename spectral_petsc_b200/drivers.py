"""The reference's two executables as command lines over the CUDA shells:

    python -m spectral_petsc_b200.elliptic -dim 16,16,16 -exact 2 -ksp_rtol 1e-10          (elliptic.C:116-247)
    python -m spectral_petsc_b200.stokes -exact 2 -cont 4 -rheology 1 -eps 1e-4 -exponent 3 -schur_ksp_max_it 3 \\
           -vel_ksp_max_it 4 -svel_ksp_type preonly -ksp_type fgmres -dim 20,20,20           (stokes.C:114-255, README:44,55)

Same options, same order of work, same printed lines (problem header, DOF distribution, norms at the exact solution, null
space test, continuation banner, iteration count / reason / norm of error, stokes.vtk).  What runs where:
  * operators, residuals, FGMRES, the finite-difference preconditioning MATRICES: the C-ABI library on the GPU;
  * SNES (full Newton steps, solvers.newton), the saddle-point PC composition (solvers.StokesSaddlePC): host orchestration;
  * the PC built FROM the finite-difference matrix is PETSc's own in the reference (ILU(2) set in code, elliptic.C:183-184;
    hypre / LU by option, README:12-13) and out of scope here: HostPC below is a scipy stand-in on the host, selected by the
    same -pc_type / -vel_pc_type / -svel_pc_type options.  Its time is PC time, not operator time.

The flows are written against small problem adapters (GpuElliptic / GpuStokes here) so tests/test_drivers_cpu.py can run
the identical flow over the CPU oracle; the adapters in this file fail without the CUDA library and a device.
"""
import math
import sys

import numpy as np

from . import solvers

SNES_REASONS = {2: "CONVERGED_FNORM_ABS", 3: "CONVERGED_FNORM_RELATIVE", -5: "DIVERGED_MAX_IT", -3: "DIVERGED_LINEAR_SOLVE"}


class OptionsError(ValueError):
    pass


class PetscOptions:
    """The slice of the PETSc options database the drivers read: `-name value` pairs and bare `-flag`s."""

    def __init__(self, argv):
        self.kv, self.used = {}, set()
        i = 0
        while i < len(argv):
            a = argv[i]
            if not a.startswith("-") or _is_number(a):
                raise OptionsError("expected an option name, got %r" % a)
            if i + 1 < len(argv) and (not argv[i + 1].startswith("-") or _is_number(argv[i + 1])):
                self.kv[a[1:]] = argv[i + 1]
                i += 2
            else:
                self.kv[a[1:]] = None
                i += 1

    def _get(self, name):
        self.used.add(name)
        return self.kv.get(name)

    def has(self, name):  # PetscOptionsHasName
        self.used.add(name)
        return name in self.kv

    def int(self, name, default):
        v = self._get(name)
        return default if v is None else int(v)

    def real(self, name, default):
        v = self._get(name)
        return default if v is None else float(v)

    def string(self, name, default):
        v = self._get(name)
        return default if v is None else v

    def int_array(self, name, default, maxlen=10):  # PetscOptionsIntArray (elliptic.C:141)
        v = self._get(name)
        if v is None:
            return list(default)
        out = [int(t) for t in v.split(",") if t != ""]
        if not out or len(out) > maxlen:
            raise OptionsError("-%s takes 1..%d comma-separated integers" % (name, maxlen))
        return out

    def unused(self):
        return sorted(set(self.kv) - self.used)


def _is_number(s):
    try:
        float(s)
        return True
    except ValueError:
        return False


class HostPC:
    """Stand-in for PETSc's PC on a finite-difference matrix (scipy CSR on the host), selected like PETSc's:
    -pc_type ilu (level-of-fill ILU(k), -pc_factor_levels k: the C++ sb200_host_ilu_*, what the reference sets in code with
    k = 2, elliptic.C:183-184) | lu (SuperLU) | ilut (scipy's threshold ILU) | jacobi | none; `hypre` (README:12) has no
    counterpart here and maps to lu.  apply() takes and returns HOST arrays; update(P) refactors for a new matrix with
    the same pattern (SAME_NONZERO_PATTERN)."""

    TYPES = ("ilu", "lu", "ilut", "jacobi", "none", "hypre")

    def __init__(self, P, pc_type="ilu", levels=0):
        if pc_type not in self.TYPES:
            raise OptionsError("unknown PC type %r (have: %s)" % (pc_type, ", ".join(self.TYPES)))
        self.type = "lu" if pc_type == "hypre" else pc_type
        self.levels = levels
        self._ilu = None
        self.update(P)

    def update(self, P):
        import scipy.sparse.linalg as spla

        if self.type == "ilu":
            if self._ilu is None:
                from .capi import HostILU

                self._ilu = HostILU(P, self.levels)
            else:
                self._ilu.refactor(P)
            self.apply = self._ilu.solve
        elif self.type == "lu":
            self.apply = spla.splu(P.tocsc()).solve
        elif self.type == "ilut":
            self.apply = spla.spilu(P.tocsc(), drop_tol=1e-4, fill_factor=10).solve
        elif self.type == "jacobi":
            dinv = 1.0 / P.diagonal()
            self.apply = lambda r: dinv * r
        else:
            self.apply = lambda r: r.copy()
        return self


# ---- problem adapters over the C-ABI library -------------------------------------------------------------------------
class GpuElliptic:
    """MatCreate_Elliptic + the callbacks, vectors as fp64 CUDA tensors."""

    def __init__(self, dim, gamma, exponent):
        import torch

        from .capi import Elliptic

        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: the drivers run the operators on the GPU only (there is no CPU fallback)")
        self.torch, self.dev = torch, torch.device("cuda", torch.cuda.current_device())
        self.G = Elliptic(dim, gamma=gamma, exponent=exponent)
        self.m, self.g, self.nd = self.G.m, self.G.g, self.G.nd
        self.krylov = solvers.make_gpu_krylov()
        self._pattern = None

    def from_host(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.dev)

    def to_host(self, v):
        return v.cpu().numpy()

    def set_dirichlet(self, a):
        self.G.set_dirichlet(self.from_host(a))

    def set_rhs(self, a):
        self.G.set_rhs(self.from_host(a))

    def form_function(self, x):
        return self.G.form_function(x).clone()

    def mat_mult(self, x):
        return self.G.mat_mult(x)

    def jacobian(self):
        """FormJacobian on the device (sb200_elliptic_jacobian_csr), handed to the host PC as scipy CSR."""
        import scipy.sparse as sps

        rowptr, colidx, vals = self.G.jacobian_csr(self._pattern)
        if self._pattern is None:
            self._pattern = (rowptr, colidx)
            self._host_pattern = (colidx.cpu().numpy(), rowptr.cpu().numpy())
        return sps.csr_matrix((vals.cpu().numpy(), self._host_pattern[0], self._host_pattern[1]), shape=(self.g, self.g))


class GpuStokes:
    """StokesCreate + the shells, vectors as fp64 CUDA tensors."""

    def __init__(self, dim, rheology, hardness, exponent, regularization, gamma0):
        import torch

        from .capi import Stokes

        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: the drivers run the operators on the GPU only (there is no CPU fallback)")
        self.torch, self.dev = torch, torch.device("cuda", torch.cuda.current_device())
        self.S = S = Stokes(dim, rheology=rheology, hardness=hardness, exponent=exponent, regularization=regularization, gamma0=gamma0)
        self.rheology, self.hardness, self.gamma0 = rheology, hardness, gamma0
        self.d, self.dim = len(dim), list(dim)
        self.m, self.g, self.gp, self.gv, self.dv = S.m, S.g, S.gp, S.gv, S.dv
        self.krylov = solvers.make_gpu_krylov()
        self._pattern = None
        # the shells, by the names solvers.StokesSaddlePC uses
        self.mat_mult, self.mat_mult_vv, self.mat_mult_pv, self.mat_mult_vp = S.mat_mult, S.mat_mult_vv, S.mat_mult_pv, S.mat_mult_vp
        self.get_diagonal_schur = S.get_diagonal_schur

    def from_host(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.dev)

    def to_host(self, v):
        return v.cpu().numpy()

    def set_dirichlet(self, a):
        self.S.set_dirichlet(self.from_host(a))

    def set_force(self, a):
        self.S.set_force(self.from_host(a))

    def set_rheology(self, exponent, regularization):
        self.S.set_rheology(self.rheology, self.hardness, exponent, regularization, self.gamma0)

    def function(self, x):
        return self.S.function(x).clone()

    def eta_minmax(self):
        return self.S.eta_minmax()

    def pc_velocity_matrix(self):
        """StokesPCSetUp0 on the device (sb200_stokes_pc_velocity_csr), handed to the host PC as scipy CSR."""
        import scipy.sparse as sps

        rowptr, colidx, vals = self.S.pc_velocity_csr(self._pattern)
        if self._pattern is None:
            self._pattern = (rowptr, colidx)
            self._host_pattern = (colidx.cpu().numpy(), rowptr.cpu().numpy())
        return sps.csr_matrix((vals.cpu().numpy(), self._host_pattern[0], self._host_pattern[1]), shape=(self.gv, self.gv))

    def state_host(self):
        """eta, deta (m) and strain[j] (m x d) of the last residual evaluation, on the host (for StokesStateView)."""
        h = self.to_host
        return h(self.S.get_state(0)), h(self.S.get_state(1)), [h(self.S.get_state(2 + j)).reshape(self.m, self.d) for j in range(self.d)]

    def pressure_reduce_order_host(self, pL):
        return self.to_host(self.S.pressure_reduce_order(self.from_host(pL)))


def _norm_inf(prob, v):
    return float(np.abs(prob.to_host(v)).max())


def _snes(prob, function, solve_jacobian, x0, rtol, atol, max_it, out, monitor):
    x, its, kits, hist = solvers.newton(function, solve_jacobian, x0, rtol=rtol, atol=atol, max_it=max_it)
    if monitor:
        for i, f in enumerate(hist):
            out("  %d SNES Function norm %.12e" % (i, f))
    if hist[-1] <= atol:
        reason = 2
    elif hist[-1] <= rtol * hist[0]:
        reason = 3
    else:
        reason = -5
    return x, its, kits, hist, reason


# ---- elliptic.C main -------------------------------------------------------------------------------------------------
def elliptic_main(argv, out=print, make_problem=GpuElliptic):
    """main() of elliptic.C (:116-247).  Returns a dict with what was printed, for tests."""
    from .capi import elliptic_exact_solution

    o = PetscOptions(argv)
    dim = o.int_array("dim", [8, 6])  # elliptic.C:139-142
    o.int("debug", 0)
    exact = o.int("exact", 0)
    gamma = o.real("gamma", 0.0)
    exponent = o.real("exponent", 2.0)
    cos_scale = o.real("cos_scale", None)
    if exact in (0, 3) and cos_scale is None:
        raise OptionsError("-exact %d needs -cos_scale (the reference reads it without a default, elliptic.C:607-609)" % exact)
    ksp_rtol, ksp_max_it = o.real("ksp_rtol", 1e-5), o.int("ksp_max_it", 10000)  # PETSc defaults
    restart = o.int("ksp_gmres_restart", 30)
    snes_rtol, snes_atol, snes_max_it = o.real("snes_rtol", 1e-8), o.real("snes_atol", 1e-50), o.int("snes_max_it", 50)
    pc_type = o.string("pc_type", "ilu")  # PCSetType(pc, PCILU); PCFactorSetLevels(pc, 2) (elliptic.C:183-184)
    pc_levels = o.int("pc_factor_levels", 2)
    ksp_type = o.string("ksp_type", "fgmres")  # KSPSetType(ksp, KSPFGMRES), elliptic.C:182
    if ksp_type != "fgmres":
        raise OptionsError("-ksp_type %s: only fgmres (the type the reference sets in code) is built" % ksp_type)
    ksp_monitor, snes_monitor = o.has("ksp_monitor"), o.has("snes_monitor")

    out("Elliptic problem  dims = [%s]    gamma = %f    exponent = %8f" % (",".join(str(v) for v in dim), gamma, exponent))
    prob = make_problem(dim, gamma, exponent)
    out("DOF distribution: %8d local     %8d global     %8d dirichlet" % (prob.m, prob.g, prob.nd))  # elliptic.C:424
    u, u2, dirichlet = elliptic_exact_solution(dim, exact, cos_scale or 0.0, gamma, exponent)  # CreateExactSolution, :187
    prob.set_dirichlet(dirichlet)
    prob.set_rhs(u2)  # VecCopy(u2, ac->b), :674
    res = {"dim": dim, "g": prob.g}

    r = prob.to_host(prob.form_function(prob.from_host(u)))  # CHECK_EXACT block, :192-209
    with np.errstate(divide="ignore", invalid="ignore"):
        res["exact_residual_abs"], res["exact_residual_rel"] = float(np.abs(r).max()), float(np.nanmax(np.abs(r / u2)))
    out("%-25s: abs = %8e   rel = %8e" % ("Norm of exact residual", res["exact_residual_abs"], res["exact_residual_rel"]))

    ksp_log, pcs = [], []

    def solve_jacobian(rhs):  # one KSPSolve of the SNES: FormJacobian -> PC set-up -> FGMRES on MatMult_Elliptic
        P = prob.jacobian()
        pc = pcs[0].update(P) if pcs else HostPC(P, pc_type, pc_levels)
        pcs[:] = [pc]
        x, its, reason = prob.krylov(prob.mat_mult, rhs, lambda v: prob.from_host(pc.apply(prob.to_host(v))), ksp_rtol, ksp_max_it, restart)
        ksp_log.append((its, reason))
        if ksp_monitor:
            out("    KSP iterations %d reason %d" % (its, reason))
        return x, its

    x0 = prob.from_host(np.zeros(prob.g))  # VecSet(x, 0.0), :212
    x, its, kits, hist, reason = _snes(prob, prob.form_function, solve_jacobian, x0, snes_rtol, snes_atol, snes_max_it, out, snes_monitor)
    e = prob.to_host(x) - u
    with np.errstate(divide="ignore", invalid="ignore"):
        res["error_abs"], res["error_rel"] = float(np.abs(e).max()), float(np.nanmax(np.abs(e / u)))
    res.update(snes_its=its, ksp_its=kits, ksp_reasons=[q for _, q in ksp_log], reason=SNES_REASONS[reason], fnorm=hist)
    out("Number of nonlinear iterations = %d" % its)
    out("Reason for solver termination: %s" % SNES_REASONS[reason])
    out("%-25s: abs = %8e   rel = %8e" % ("Norm of error", res["error_abs"], res["error_rel"]))
    for name in o.unused():
        out("WARNING! There are options you set that were not used: -%s" % name)  # PETSc's own message at PetscFinalize
    res["x"] = prob.to_host(x)
    return res


# ---- stokes.C main ---------------------------------------------------------------------------------------------------
def _pad_local(dim, interior_vals, boundary_vals, ncomp):
    """scatterVL/scatterPL + scatterDL on the host: full-grid array [m, ncomp] from interior values (walk order) and
    boundary values (walk order; None = zeros)."""
    idx = np.indices(dim).reshape(len(dim), -1)
    on_bdy = np.zeros(idx.shape[1], dtype=bool)
    for j, n in enumerate(dim):
        on_bdy |= (idx[j] == 0) | (idx[j] == n - 1)
    L = np.zeros((idx.shape[1], ncomp))
    L[~on_bdy] = interior_vals.reshape(-1, ncomp)
    if boundary_vals is not None:
        L[on_bdy] = boundary_vals.reshape(-1, ncomp)
    return L


def write_stokes_vtk(path, dim, coord, vel, pres, vel_force, div_force, eta, deta, strain):
    """StokesStateView (stokes.C:1821-1894): legacy ASCII VTK structured grid with the reference's fields and number format."""
    d = len(dim)
    m, n, p = dim[0], dim[1], (dim[2] if d > 2 else 1)
    nodes = int(np.prod(dim))

    def vec_view(f, a, pernode, perline):  # StokesVecView (stokes.C:1898-1915)
        a = a.reshape(nodes, pernode)
        for i in range(nodes):
            f.write("".join("%20e " % a[i, j] for j in range(min(pernode, perline))) + "0 " * max(perline - pernode, 0) + "\n")

    with open(path, "w") as f:
        f.write("# vtk DataFile Version 2.0\nStokes Output\nASCII\nDATASET STRUCTURED_GRID\n")
        f.write("DIMENSIONS %d %d %d\nPOINTS %d double\n" % (m, n, p, m * n * p))
        vec_view(f, coord, d, 3)
        f.write("\nPOINT_DATA %d\nVECTORS velocity double\n" % (m * n * p))
        vec_view(f, vel, d, 3)
        f.write("\nSCALARS pressure double 1\nLOOKUP_TABLE default\n")
        vec_view(f, pres, 1, 1)
        f.write("\nVECTORS vel_force double\n")
        vec_view(f, vel_force, d, 3)
        f.write("\nSCALARS div_force double 1\nLOOKUP_TABLE default\n")
        vec_view(f, div_force, 1, 1)
        f.write("\nSCALARS eta double 1\nLOOKUP_TABLE default\n")
        vec_view(f, eta, 1, 1)
        f.write("\nSCALARS deta double 1\nLOOKUP_TABLE default\n")
        vec_view(f, deta, 1, 1)
        f.write("\nTENSORS strain double\n")
        for i in range(nodes):
            for j in range(3):
                f.write("".join("%20e " % (strain[j][i, k] if (j < d and k < d) else 0.0) for k in range(3)) + "\n")
            f.write("\n")


def stokes_state_view(prob, x_host, force_host, dirichlet_host, path="stokes.vtk"):
    """The scatters of StokesStateView (stokes.C:1827-1850) on the host, then the writer.  The pressure fields are extended
    to the boundary nodes by StokesPressureReduceOrder, the velocities get the Dirichlet values."""
    d, dim = prob.d, prob.dim
    idx = np.indices(dim).reshape(d, -1)
    coord = np.stack([np.cos(idx[j] * math.pi / (dim[j] - 1)) for j in range(d)], axis=1)
    fields = []
    for g in (x_host, force_host):
        v, p = solvers.split(g, d)
        fields.append(_pad_local(dim, v, dirichlet_host, d))
        fields.append(prob.pressure_reduce_order_host(_pad_local(dim, p, None, 1).reshape(-1)))
    eta, deta, strain = prob.state_host()
    write_stokes_vtk(path, dim, coord, fields[0], fields[1], fields[2], fields[3], eta, deta, strain)
    return path


def stokes_main(argv, out=print, make_problem=GpuStokes):
    """main() of stokes.C (:114-255) with StokesProcessOptions (:392-495), -boundary 0 (all Dirichlet).  Returns a dict."""
    from .capi import stokes_exact_solution

    o = PetscOptions(argv)
    dim = o.int_array("dim", [8, 6])  # stokes.C:407-408
    o.int("debug", 0)
    exact, boundary, rheology = o.int("exact", 0), o.int("boundary", 0), o.int("rheology", 0)
    hardness, exponent = o.real("hardness", 1.0), o.real("exponent", 1.0)
    regularization, gamma0 = o.real("eps", 1.0), o.real("gamma0", 1.0)
    cont0, cont = o.int("cont0", 0), o.int("cont", 1)
    for name in ("scaleM", "scaleN", "zeroV"):  # mixed / Neumann boundary knobs: read like the reference, unused with -boundary 0
        o.real(name, 1.0)
    o.int("zeroN", 0)
    if len(dim) not in (2, 3):
        raise OptionsError("the Stokes driver needs 2 or 3 dimensions (StokesPressureReduceOrder, stokes.C:1036)")
    if boundary != 0:
        raise OptionsError("Boundary type %d not implemented (README:64-68: the Neumann / mixed conditions are broken upstream)" % boundary)
    if rheology not in (0, 1):
        raise OptionsError("Rheology type %d not implemented" % rheology)  # stokes.C:492
    pcvel, saddle = o.int("pcvel", 0), o.int("pc_saddle_type", 0)
    if pcvel != 0:
        raise OptionsError("pcvel type number %d not implemented (only the finite-difference matrix, StokesPCSetUp0)" % pcvel)
    if saddle not in (0, 1, 2, 3):
        raise OptionsError("pc_saddle_type %d not implemented" % saddle)  # stokes.C:184
    ksp_type = o.string("ksp_type", "fgmres")
    if ksp_type != "fgmres":
        raise OptionsError("-ksp_type %s: only fgmres (the type the reference sets in code, stokes.C:157) is built" % ksp_type)
    ksp_rtol, ksp_max_it = o.real("ksp_rtol", 1e-5), o.int("ksp_max_it", 10000)
    snes_rtol, snes_atol, snes_max_it = o.real("snes_rtol", 1e-8), o.real("snes_atol", 1e-50), o.int("snes_max_it", 50)
    vel_max_it, vel_rtol = o.int("vel_ksp_max_it", 10000), o.real("vel_ksp_rtol", 1e-5)
    schur_max_it, schur_rtol = o.int("schur_ksp_max_it", 10000), o.real("schur_ksp_rtol", 1e-5)
    svel_preonly = o.string("svel_ksp_type", "gmres") == "preonly"
    svel_max_it, svel_rtol = o.int("svel_ksp_max_it", 10000), o.real("svel_ksp_rtol", 1e-5)  # KSPSchurVelocity's own prefix (stokes.C:338-341)
    # PETSc's default PC for the SeqAIJ matrix MatVVPC is ILU(0); README:44 overrides it with hypre
    vel_pc, svel_pc = o.string("vel_pc_type", "ilu"), o.string("svel_pc_type", "ilu")
    vel_levels, svel_levels = o.int("vel_pc_factor_levels", 0), o.int("svel_pc_factor_levels", 0)
    ksp_monitor, snes_monitor = o.has("ksp_monitor"), o.has("snes_monitor")
    want_vtk = o.has("output_vtk")
    vtk_path = o.string("output_vtk", None) or "stokes.vtk"

    d = len(dim)
    out("Stokes problem  dim = [%s]" % ",".join(str(v) for v in dim))
    out("  hardness = %f    exponent = %8f    regularization = %8f    gamma0 = %8f" % (hardness, exponent, regularization, gamma0))
    prob = make_problem(dim, rheology, hardness, exponent, regularization, gamma0)
    out("DOF distribution: %d global   %d/%d pressure    %d/%d velocity    %d dirichlet    %d mixed"
        % (prob.g, prob.gp, prob.m, prob.gv, prob.m * d, prob.dv, 0))  # stokes.C:891
    U, U2, dirichlet = stokes_exact_solution(dim, exact)  # StokesCreateExactSolution, :178
    prob.set_dirichlet(dirichlet)
    prob.set_force(U2)  # VecCopy(U2, c->force), :1001
    res = {"dim": dim, "g": prob.g, "steps": []}

    def function(x):
        F = prob.function(x)
        mn, mx = prob.eta_minmax()
        out("Minimum eta = %9.3e   Maximum eta = %9.3e" % (mn, mx))  # stokes.C:731-734, printed by every residual evaluation
        return F

    r = function(prob.from_host(U))
    res["exact_residual"] = _norm_inf(prob, r)
    out("Norm of solution %9.3e  norm of forcing %9.3e  norm of residual %9.3e" % (float(np.abs(U).max()), float(np.abs(U2).max()), res["exact_residual"]))
    ns = np.zeros(prob.g)
    ns[d::d + 1] = 1.0 / math.sqrt(prob.gp)  # StokesRemoveConstantPressure: the normalised constant-pressure vector, :1013-1020
    res["null_space"] = _norm_inf(prob, prob.mat_mult(prob.from_host(ns)))
    if not res["null_space"] < 1e-8:  # MatNullSpaceTest, :206-212
        raise RuntimeError("Null space test failed")

    pcs = {}

    def make_saddle_pc():  # StokesPCSetUp0 + the PCs PETSc builds on MatVVPC for KSPVelocity / KSPSchurVelocity
        P = prob.pc_velocity_matrix()
        pcs["vel"] = pcs["vel"].update(P) if "vel" in pcs else HostPC(P, vel_pc, vel_levels)
        if (svel_pc, svel_levels) == (vel_pc, vel_levels):
            pcs["svel"] = pcs["vel"]
        else:
            pcs["svel"] = pcs["svel"].update(P) if "svel" in pcs else HostPC(P, svel_pc, svel_levels)
        on_dev = lambda pc: (lambda v: prob.from_host(pc.apply(prob.to_host(v))))
        spc = solvers.StokesSaddlePC(prob, d, prob.krylov, on_dev(pcs["vel"]), saddle_type=saddle, vel_max_it=vel_max_it, schur_max_it=schur_max_it,
                                     vel_rtol=vel_rtol, schur_rtol=schur_rtol, svel_preonly=svel_preonly, svel_rtol=svel_rtol, svel_max_it=svel_max_it)
        if pcs["svel"] is not pcs["vel"]:
            svel = on_dev(pcs["svel"])
            if svel_preonly:
                spc.solve_schur_velocity = svel
            else:
                spc.solve_schur_velocity = lambda rhs: solvers.left_gmres(prob.krylov, prob.mat_mult_vv, svel, rhs, svel_rtol, svel_max_it)[0]
        return spc

    x = prob.from_host(np.zeros(prob.g))  # VecSet(x, 0.0), :215
    for i in range(cont0, cont + 1):  # the continuation loop, :216-236
        e_i, r_i = solvers.continuation_params(i, cont, exponent, regularization)
        prob.set_rheology(e_i, r_i)
        out("## [%d/%d] Solving with exponent = %5f regularization %8.2e" % (i, cont, e_i, r_i))

        def solve_jacobian(rhs):
            dx, its, reason = solvers.solve_stokes_linear(prob, d, prob.krylov, make_saddle_pc(), rhs, rtol=ksp_rtol, maxits=ksp_max_it)
            if ksp_monitor:
                out("    KSP iterations %d reason %d" % (its, reason))
            return dx, its

        x, its, kits, hist, reason = _snes(prob, function, solve_jacobian, x, snes_rtol, snes_atol, snes_max_it, out, snes_monitor)
        err = solvers.remove_constant_pressure(prob.to_host(x) - U, d)  # VecAXPY(r, -1, u); MatNullSpaceRemove, :224-226
        step = {"step": i, "exponent": e_i, "regularization": r_i, "snes_its": its, "ksp_its": kits, "reason": SNES_REASONS[reason],
                "error": float(np.abs(err).max()), "fnorm": hist}
        res["steps"].append(step)
        out("Number of nonlinear iterations = %d" % its)
        out("Reason for solver termination: %s" % SNES_REASONS[reason])
        out("%-25s: abs = %8e" % ("Norm of error", step["error"]))
    if want_vtk:  # StokesStateView(ctx, x, "final state"), :238-242
        res["vtk"] = stokes_state_view(prob, prob.to_host(x), U2, dirichlet, vtk_path)
    for name in o.unused():
        out("WARNING! There are options you set that were not used: -%s" % name)
    res["x"] = prob.to_host(x)
    return res


def _run(main, argv):
    try:
        main(argv)
    except OptionsError as e:
        print("error: %s" % e, file=sys.stderr)
        return 83  # PETSC_ERR_USER
    return 0
