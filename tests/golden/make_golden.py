"""Regenerates tests/golden/*.npz from the oracle (run from the repo root: python tests/golden/make_golden.py).

These vectors are outputs of the ORACLE restatement on seeded inputs.  tests/test_oracle_ref*.py check them against the
reference's own chebyshev.c / elliptic.C / stokes.C (compiled against FFTW / PETSc stand-ins, oracle/_ref):
the cheb, elliptic and stokes vectors here equal what the reference source produces to 1e-12 or better.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.chebyshev import ChebCtx, cheb_mult  # noqa: E402
from oracle.elliptic import MatElliptic  # noqa: E402
from oracle.stokes import StokesCtx  # noqa: E402
from oracle.fgmres import fgmres  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def cheb_case():
    dims = [8, 7, 6]  # cheb.c:77-91 (K2): u = e^x + e^y + e^z on the CGL grid
    idx = np.indices(dims).reshape(3, -1)
    x = [np.cos(idx[j] * np.pi / (dims[j] - 1)) for j in range(3)]
    u = np.exp(x[0]) + np.exp(x[1]) + np.exp(x[2])
    rnd = np.random.default_rng(7).standard_normal(u.size)
    out = {"dims": np.array(dims), "u": u, "rnd": rnd}
    for tr in range(3):
        out["du%d" % tr] = cheb_mult(ChebCtx(3, tr, dims), u)
        out["drnd%d" % tr] = cheb_mult(ChebCtx(3, tr, dims), rnd)
    np.savez_compressed(os.path.join(OUT, "cheb_8x7x6.npz"), **out)


def elliptic_case(dim, name):
    O = MatElliptic(dim, gamma=4.0, exponent=2.0)
    O.create_exact_solution(2)
    Us = 0.1 * np.random.default_rng(1).standard_normal(O.g)
    U = np.random.default_rng(0).standard_normal(O.g)
    F = O.form_function(Us)
    V = O.mat_mult(U)
    np.savez_compressed(os.path.join(OUT, name), dim=np.array(dim), dirichlet=O.dirichlet, b=O.b, Us=Us, U=U, F=F, V=V, eta=O.eta,
                        gradu0=O.gradu[0])


def stokes_case(dim, name):
    O = StokesCtx(dim, rheology=1, hardness=1.0, exponent=3.0, regularization=1e-2, gamma0=1.0, exact=2)
    O.create_exact_solution()
    xs = 0.3 * np.random.default_rng(1).standard_normal(O.g)
    x = np.random.default_rng(0).standard_normal(O.g)
    F = O.function(xs)
    y = O.mat_mult(x)
    np.savez_compressed(os.path.join(OUT, name), dim=np.array(dim), dirichlet=O.dirichlet.reshape(-1), force=O.force, xs=xs, x=x, F=F, y=y,
                        eta=O.eta)


def ksp_case():
    import scipy.sparse.linalg as spla

    dim = [8, 8, 8]
    O = MatElliptic(dim, gamma=0.0)
    u, _ = O.create_exact_solution(2)
    F0 = O.form_function(np.zeros(O.g))
    lu = spla.splu(O.form_jacobian_matrix().tocsc())
    dx, its, hist, reason = fgmres(O.mat_mult, -F0, M=lu.solve, rtol=1e-10)
    np.savez_compressed(os.path.join(OUT, "ksp_elliptic_8x8x8_exact2.npz"), dim=np.array(dim), its=its, hist=np.array(hist), dx=dx, u=u)


def saddle_case():
    """StokesPCApply0..3 (stokes.C:1714-1817) over the oracle shells and the oracle FGMRES, by spectral_petsc_b200.solvers (the
    composition tests/test_gpu_solvers.py pins to the reference's flow): power-law state of the manufactured solution, Jacobi on
    MatVVPC as the velocity PC, -vel_ksp_max_it 4 -schur_ksp_max_it 3 -svel_ksp_type preonly."""
    from spectral_petsc_b200 import solvers

    dim = [7, 6, 5]
    O = StokesCtx(dim, rheology=1, exponent=2.0, regularization=0.5, exact=2)
    U, _ = O.create_exact_solution()
    O.function(U)
    dinv = 1.0 / O.pc_velocity_matrix().diagonal()

    def krylov(op, b, pc, rtol, maxits, restart):
        x, its, hist, reason = fgmres(op, b, M=pc, restart=restart, rtol=rtol, maxits=maxits)
        return x, its, reason

    x = np.random.default_rng(11).standard_normal(O.g)
    out = {"dim": np.array(dim), "x": x}
    for t in range(4):
        pc = solvers.StokesSaddlePC(O, 3, krylov, lambda r: dinv * r, saddle_type=t, vel_max_it=4, schur_max_it=3, svel_preonly=True)
        out["y%d" % t] = pc.apply(x)
        out["its%d" % t] = np.array([pc.inner_its["velocity"], pc.inner_its["schur"]])
    np.savez_compressed(os.path.join(OUT, "saddle_7x6x5.npz"), **out)


if __name__ == "__main__":
    saddle_case()
    cheb_case()
    elliptic_case([8, 6], "elliptic_8x6.npz")
    elliptic_case([7, 6, 5], "elliptic_7x6x5.npz")
    elliptic_case([16, 16, 16], "elliptic_16x16x16.npz")
    stokes_case([8, 6], "stokes_8x6.npz")
    stokes_case([9, 7, 6], "stokes_9x7x6.npz")
    ksp_case()
    print("golden fixtures written to", OUT)
