"""Runs the C++ driver that talks to the B200 path ONLY through the reference's own interface
(MatCreateCheb / ChebMult / MatCreate_Elliptic / FormFunction / StokesCreate / StokesFunction ...)
and checks the reference's self-check outputs (cheb.c, elliptic.C:193-209, stokes.C:190-212)."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_api_driver(cuda):
    exe = os.path.join(ROOT, "tests", "cpp", "ref_api_driver")
    assert os.path.exists(exe), "run `make` (or __graft_entry__.build()) first"
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    txt = out.stdout
    f = lambda pat: float(re.search(pat, txt).group(1))
    assert f(r"cheb1d Norm of error (\S+)") == pytest.approx(1.0293308609854e-02, rel=1e-9)       # cheb.c, m1 = 5
    assert f(r"cheb3d axis 0 Norm of error (\S+)") == pytest.approx(6.245e-06, rel=5e-3)
    assert f(r"cheb3d axis 1 Norm of error (\S+)") == pytest.approx(8.72e-05, rel=5e-3)
    assert f(r"cheb3d axis 2 Norm of error (\S+)") == pytest.approx(1.04e-03, rel=5e-3)
    assert "cheb bad tr -> 83" in txt                                                           # chebyshev.c:106
    res = [float(x) for x in re.findall(r"Norm of exact residual\s*: abs = (\S+)", txt)]
    assert len(res) == 2 and res[0] < 5e-11 and res[1] < 5e-11                                  # 16^3 and 12^5, -exact 2
    assert "elliptic global dofs 2744" in txt and "elliptic global dofs 100000" in txt          # elliptic.C:424
    assert f(r"norm of residual\s+(\S+)") < 2e-11                                               # stokes 20^3 -exact 2
    assert f(r"Norm of solution\s+(\S+)") == pytest.approx(0.991, rel=1e-3)
    assert f(r"Null space test \|A ns\| =\s+(\S+)") < 1e-12                                     # stokes.C:206-212
