"""The FGMRES oracle (oracle/fgmres.py) against known answers: exact convergence on small SPD / nonsymmetric
systems, agreement with scipy's GMRES solution, and the reference's own solver-level check - "Norm of error"
after `./elliptic -dim 16,16,16 -exact 2 -ksp_rtol 1e-10` (elliptic.C:207-226, BASELINE config 1) ~1e-11."""
import numpy as np
import scipy.sparse.linalg as spla

from oracle.elliptic import MatElliptic
from oracle.fgmres import fgmres


def test_small_systems():
    rng = np.random.default_rng(0)
    n = 40
    A = rng.standard_normal((n, n)) + n * np.eye(n)
    b = rng.standard_normal(n)
    x, its, hist, reason = fgmres(lambda v: A @ v, b, rtol=1e-12)
    assert reason == 2 and its <= n
    assert np.linalg.norm(A @ x - b) <= 1e-11 * np.linalg.norm(b)
    assert np.all(np.diff(hist) <= 1e-14)  # GMRES residuals never increase
    # restart path
    x2, its2, _, reason2 = fgmres(lambda v: A @ v, b, restart=5, rtol=1e-10)
    assert reason2 == 2 and np.linalg.norm(A @ x2 - b) <= 1e-9 * np.linalg.norm(b)
    # exact PC: one iteration
    x3, its3, _, _ = fgmres(lambda v: A @ v, b, M=lambda v: np.linalg.solve(A, v), rtol=1e-12)
    assert its3 == 1 and np.allclose(x3, np.linalg.solve(A, b))


def test_config1_elliptic_16_exact2_solver_level():
    O = MatElliptic([16, 16, 16], gamma=0.0)
    u, _ = O.create_exact_solution(2)
    F0 = O.form_function(np.zeros(O.g))  # also sets the (trivial) state
    lu = spla.splu(O.form_jacobian_matrix().tocsc())
    dx, its, hist, reason = fgmres(O.mat_mult, -F0, M=lu.solve, rtol=1e-10)
    assert reason == 2 and 5 <= its <= 60
    err = np.abs(dx - u).max()
    assert err < 1e-9  # "Norm of error" of the reference run is ~1e-11
