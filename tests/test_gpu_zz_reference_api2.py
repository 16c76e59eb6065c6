"""(Runs late in the GPU suite.)  Second C++ driver over the reference's own interface: MatCreateChebD1 / ChebD1Mult
(chebyshev.c:8-85) and the Schur shell StokesMatMultSchur with its inner KSP registered as a callback (stokes.C:318, 523-535)."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_api_driver2(cuda):
    exe = os.path.join(ROOT, "tests", "cpp", "ref_api_driver2")
    assert os.path.exists(exe), "run `make` (or __graft_entry__.build()) first"
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    txt = out.stdout
    assert "chebD1 vs cheb max diff 0.000e+00" in txt and "chebD1 n=1 -> 83" in txt  # same operator; "n = 1 but must be >= 2" (chebyshev.c:18)
    assert "Schur without an inner solve -> 62" in txt
    m = re.search(r"Schur identity-solve calls (\d+)  max \|S p \+ PV VP p\| / max \|PV VP p\| = (\S+)", txt)
    assert int(m.group(1)) == 1 and float(m.group(2)) < 1e-14
