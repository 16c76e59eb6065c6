"""(Runs late in the GPU suite.)  Second C++ driver over the reference's own interface: MatCreateChebD1 / ChebD1Mult
(chebyshev.c:8-85) and the Schur shell StokesMatMultSchur with its inner KSP registered as a callback (stokes.C:318, 523-535)."""
import os
import subprocess

import pytest

from support.ref_api_checks import check_driver2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_api_driver2(cuda):
    exe = os.path.join(ROOT, "tests", "cpp", "ref_api_driver2")
    assert os.path.exists(exe), "run `make` (or __graft_entry__.build()) first"
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    check_driver2(out.stdout)
