"""Oracle restatement of PETSc's KSPFGMRES as the reference uses it (test infrastructure).

The reference selects the solver in code - KSPSetType(ksp, KSPFGMRES) at elliptic.C:181-182 and stokes.C:155-157 -
and leaves everything else at PETSc's defaults; PETSc itself (unpinned, ~3.0.0 API, README:4-5) is not under
/root/reference, so this follows its published algorithm (src/ksp/ksp/impls/gmres/fgmres/fgmres.c):
  restart 30, right (flexible) preconditioning, classical Gram-Schmidt without refinement, Givens-rotated Hessenberg,
  residual norm from the recurrence, KSPConvergedDefault: converged when rnorm <= max(rtol*||b||, atol).
Parity unpinned beyond the reference's own analytic checks ("Norm of error" after the solve, elliptic.C:207-226).
"""
import numpy as np


def fgmres(A, b, M=None, x0=None, restart=30, rtol=1e-5, atol=1e-50, dtol=1e5, maxits=10000):
    """Returns (x, its, history, reason); A and M are callables on numpy vectors."""
    n = b.size
    x = np.zeros(n) if x0 is None else x0.copy()
    bnorm = np.linalg.norm(b)
    ttol = max(rtol * bnorm, atol)
    its, hist, reason = 0, [], 0
    while True:
        r = b - A(x) if (x0 is not None or its > 0) else b.copy()
        beta = np.linalg.norm(r)
        if its == 0:
            hist.append(beta)
        if beta <= ttol:
            return x, its, hist, 2
        V = np.zeros((restart + 1, n))
        Z = np.zeros((restart, n))
        H = np.zeros((restart + 1, restart))
        cs, sn, g = np.zeros(restart), np.zeros(restart), np.zeros(restart + 1)
        V[0] = r / beta
        g[0] = beta
        k, done = 0, False
        while k < restart and not done:
            Z[k] = V[k] if M is None else M(V[k])
            w = A(Z[k])
            h = V[:k + 1] @ w  # classical Gram-Schmidt: all projections from the same w
            w = w - h @ V[:k + 1]
            hn = np.linalg.norm(w)
            H[:k + 1, k] = h
            for i in range(k):
                a, c = H[i, k], H[i + 1, k]
                H[i, k] = cs[i] * a + sn[i] * c
                H[i + 1, k] = -sn[i] * a + cs[i] * c
            a = H[k, k]
            rr = np.hypot(a, hn)
            cs[k], sn[k] = (a / rr, hn / rr) if rr > 0 else (1.0, 0.0)
            H[k, k] = rr
            g[k + 1] = -sn[k] * g[k]
            g[k] = cs[k] * g[k]
            V[k + 1] = w / hn if hn > 0 else 0.0
            rn = abs(g[k + 1])
            its += 1
            k += 1
            hist.append(rn)
            if rn <= ttol:
                reason, done = 2, True
            elif rn >= dtol * bnorm:
                reason, done = -4, True
            elif its >= maxits:
                reason, done = -3, True
        y = np.linalg.solve(np.triu(H[:k, :k]), g[:k]) if k else np.zeros(0)
        x = x + y @ Z[:k]
        if done:
            return x, its, hist, reason
