"""Host-side orchestration of the reference's solver stack around the shells: pure composition of
operator applications and Krylov solves, written against array objects that numpy and torch both provide, so the same code
drives the CUDA shells (torch tensors + the device FGMRES) and, in the tests, the CPU oracle.

  StokesSaddlePC   StokesPCApply0..3 (stokes.C:1714-1817): block LU / upper / diagonal / lower saddle-point preconditioners
                   built from KSPVelocity, KSPSchur and KSPSchurVelocity (stokes.C:328-341) and the PV / VP / Schur shells
  left_gmres       PETSc's default KSP for the inner solves: GMRES(30) with LEFT preconditioning and a small -ksp_max_it
  solve_stokes_linear   the outer FGMRES of stokes.C:155-157 with the constant-pressure null space removed
                   (StokesRemoveConstantPressure, stokes.C:1006-1025)
  newton           the SNES loop reduced to full Newton steps with a simple backtracking safeguard

The preconditioner of the velocity block (hypre / LU on the finite-difference matrix MatVVPC, README:38-44) stays outside:
it is passed in as a callable.  `krylov(op, b, pc, rtol, maxits, restart) -> (x, its, reason)` is the Krylov engine:
spectral_petsc_b200.KSP on the GPU (make_gpu_krylov), oracle.fgmres in the CPU tests.
"""


def split(x, d):
    """Global AoS vector [v_0..v_{d-1}, p] per interior node -> (velocity, pressure) (scatterGV / scatterGP, stokes.C:867-877)."""
    n = x.shape[0] // (d + 1)
    X = x.reshape(n, d + 1)
    v = X[:, :d].reshape(-1)
    p = X[:, d].reshape(-1)
    return (v.contiguous(), p.contiguous()) if hasattr(v, "contiguous") else (v.copy(), p.copy())


def merge(v, p, d):
    """(velocity, pressure) -> global AoS vector (scatterVG / scatterPG)."""
    n = p.shape[0]
    if hasattr(v, "new_empty"):
        x = v.new_empty(n * (d + 1))
    else:
        import numpy as np

        x = np.empty(n * (d + 1))
    X = x.reshape(n, d + 1)
    X[:, :d] = v.reshape(n, d)
    X[:, d] = p
    return x


def left_gmres(krylov, A, Minv, b, rtol=1e-5, maxits=10000, restart=30, project=None):
    """KSPGMRES with left preconditioning (PETSc's default KSP type and side): GMRES on M^-1 A with the preconditioned
    residual norm, zero initial guess.  `project` (KSPSetNullSpace) is applied after every preconditioner application."""
    if project is None:
        pb = Minv(b)
        op = lambda x: Minv(A(x))
    else:
        pb = project(Minv(b))
        op = lambda x: project(Minv(A(x)))
    return krylov(op, pb, None, rtol, maxits, restart)


class StokesSaddlePC:
    """StokesPCApply0..3 (stokes.C:1714-1817).  shells: an object with mat_mult_vv / mat_mult_pv / mat_mult_vp /
    get_diagonal_schur (spectral_petsc_b200.Stokes or the oracle's StokesCtx).  velocity_pc: callable z = M^-1 r for the
    velocity block (stands for -vel_pc_type / -svel_pc_type hypre on MatVVPC)."""

    def __init__(self, shells, d, krylov, velocity_pc, saddle_type=0, vel_max_it=4, schur_max_it=3, vel_rtol=1e-5, schur_rtol=1e-5,
                 svel_preonly=True, svel_rtol=1e-5, svel_max_it=10000):
        self.s, self.d, self.krylov, self.vpc = shells, d, krylov, velocity_pc
        self.type, self.vel_max_it, self.schur_max_it = saddle_type, vel_max_it, schur_max_it
        self.vel_rtol, self.schur_rtol, self.svel_preonly = vel_rtol, schur_rtol, svel_preonly
        self.svel_rtol, self.svel_max_it = svel_rtol, svel_max_it  # KSPSchurVelocity's own "svel_" prefix (stokes.C:338-341)
        self.inner_its = {"velocity": 0, "schur": 0}

    # KSPVelocity: -vel_ksp_max_it 4, PC on MatVVPC (stokes.C:334-337)
    def solve_velocity(self, rhs):
        x, its, _ = left_gmres(self.krylov, self.s.mat_mult_vv, self.vpc, rhs, self.vel_rtol, self.vel_max_it)
        self.inner_its["velocity"] += its
        return x

    # KSPSchurVelocity: -svel_ksp_type preonly (one application of the PC) unless told otherwise (stokes.C:338-341)
    def solve_schur_velocity(self, rhs):
        if self.svel_preonly:
            return self.vpc(rhs)
        x, _, _ = left_gmres(self.krylov, self.s.mat_mult_vv, self.vpc, rhs, self.svel_rtol, self.svel_max_it)
        return x

    # StokesMatMultSchur (stokes.C:523-535): S p = -PV (A^-1 (VP p))
    def schur(self, p):
        return -1.0 * self.s.mat_mult_pv(self.solve_schur_velocity(self.s.mat_mult_vp(p)))

    # KSPSchur: Jacobi from StokesMatGetDiagonalSchur (the "diagonal" is 1/eta, so the PC multiplies by eta), -schur_ksp_max_it 3,
    # constant null space (stokes.C:328-333, 1022-1023)
    def solve_schur(self, rhs):
        diag = self.s.get_diagonal_schur()
        x, its, _ = left_gmres(self.krylov, self.schur, lambda r: r / diag, rhs, self.schur_rtol, self.schur_max_it, project=lambda q: q - q.mean())
        self.inner_its["schur"] += its
        return x

    def apply(self, x):
        d, t = self.d, self.type
        xv, xp = split(x, d)
        if t == 0:  # full block LU (stokes.C:1714-1745)
            v1 = self.solve_velocity(xv)
            p1 = self.solve_schur(xp - self.s.mat_mult_pv(v1))
            return merge(v1 + self.solve_velocity(-1.0 * self.s.mat_mult_vp(p1)), p1, d)
        if t == 1:  # block upper triangular (stokes.C:1747-1771)
            p1 = self.solve_schur(xp)
            return merge(self.solve_velocity(xv - self.s.mat_mult_vp(p1)), p1, d)
        if t == 2:  # block diagonal (stokes.C:1773-1795)
            return merge(self.solve_velocity(xv), self.solve_schur(xp), d)
        if t == 3:  # block lower triangular (stokes.C:1797-1817)
            v1 = self.solve_velocity(xv)
            return merge(v1, self.solve_schur(xp - self.s.mat_mult_pv(v1)), d)
        raise ValueError("pc_saddle_type %d not implemented" % t)  # stokes.C:184


def remove_constant_pressure(x, d):
    """MatNullSpaceRemove with the normalised [0; 1_p] vector (StokesRemoveConstantPressure, stokes.C:1006-1025)."""
    v, p = split(x, d)
    return merge(v, p - p.mean(), d)


def solve_stokes_linear(shells, d, krylov, pc, b, rtol=1e-10, maxits=10000, restart=30):
    """The outer KSP of stokes.C:155-160: FGMRES on StokesMatMult, right-preconditioned by the saddle PC, with the constant
    pressure removed after every preconditioner application (KSPSetNullSpace)."""
    return krylov(shells.mat_mult, b, lambda r: remove_constant_pressure(pc.apply(r), d), rtol, maxits, restart)


def newton(function, solve_jacobian, x0, rtol=1e-8, atol=1e-50, max_it=50, norm=None):
    """SNESSolve reduced to its skeleton: x <- x + lambda * J^-1 (-F(x)), lambda halved while the residual norm does not
    decrease (PETSc's cubic line search is PETSc's own).  function(x) must refresh the Jacobian state (as FormFunction /
    StokesFunction do); solve_jacobian(rhs) returns (dx, ksp_its).  Returns (x, newton_its, ksp_its_per_step, fnorms)."""
    if norm is None:
        norm = lambda v: float((v * v).sum() ** 0.5)
    x = x0
    F = function(x)
    f0 = fn = norm(F)
    hist, kits = [fn], []
    its = 0
    while its < max_it and fn > max(rtol * f0, atol):
        dx, k = solve_jacobian(-1.0 * F)
        kits.append(k)
        lam = 1.0
        while True:
            xn = x + lam * dx
            Fn = function(xn)
            fnn = norm(Fn)
            if fnn < fn or lam < 1e-3:
                break
            lam *= 0.5
        x, F, fn = xn, Fn, fnn
        hist.append(fn)
        its += 1
    return x, its, kits, hist


def continuation_params(i, cont, exponent, regularization):
    """Step i of the continuation loop (stokes.C:217-219): exponent_i = 1 + (i/cont)^0.8 (n - 1), eps_i = exp(log(eps) i/cont)."""
    import math

    t = 1.0 * i / cont
    return 1.0 + math.pow(t, 0.8) * (exponent - 1.0), math.exp(math.log(regularization) * i / cont)


def solve_stokes_continuation(function, set_rheology, make_saddle_pc, shells, d, krylov, x0, exponent, regularization, cont=1, cont0=0,
                              ksp_rtol=1e-5, snes_rtol=1e-8, ksp_maxits=10000, snes_max_it=50):
    """The solve loop of stokes.C:214-236: for i = cont0..cont set the rheology of step i, run SNES from the previous solution.
    function(x): the residual (refreshes the Jacobian state); set_rheology(exponent_i, eps_i); make_saddle_pc(): a
    StokesSaddlePC for the current state (the velocity-block PC matrix is assembled by the caller, PETSc's side).
    Returns (x, [per-step dict with the parameters, SNES iterations, KSP iterations per Newton step, residual norms])."""
    x, log = x0, []
    for i in range(cont0, cont + 1):
        e, r = continuation_params(i, cont, exponent, regularization)
        set_rheology(e, r)

        def solve_jacobian(rhs):
            pc = make_saddle_pc()
            dx, its, _ = solve_stokes_linear(shells, d, krylov, pc, rhs, rtol=ksp_rtol, maxits=ksp_maxits)
            return dx, its

        x, its, kits, hist = newton(function, solve_jacobian, x, rtol=snes_rtol, max_it=snes_max_it)
        log.append({"step": i, "exponent": e, "regularization": r, "snes_its": its, "ksp_its": kits, "fnorm": hist})
    return x, log


def make_gpu_krylov():
    """krylov engine over the device FGMRES (spectral_petsc_b200.KSP); one solver object per (n, restart)."""
    from .capi import KSP

    cache = {}

    def krylov(op, b, pc, rtol, maxits, restart):
        key = (int(b.numel()), int(restart))
        if key not in cache:
            cache[key] = KSP(key[0], restart=key[1])
        K = cache[key]
        K.set_operators(op, pc=pc)
        K.set_tolerances(rtol=rtol, maxits=maxits)
        x = K.solve(b).clone()
        r = K.result
        K._keep.clear()
        return x, r["its"], r["reason"]

    return krylov
