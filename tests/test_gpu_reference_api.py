"""Runs the C++ driver that talks to the B200 path ONLY through the reference's own interface
(MatCreateCheb / ChebMult / MatCreate_Elliptic / FormFunction / StokesCreate / StokesFunction ...)
and checks the reference's self-check outputs (cheb.c, elliptic.C:193-209, stokes.C:190-212)."""
import os
import subprocess

import pytest

from support.ref_api_checks import check_driver1

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_api_driver(cuda):
    exe = os.path.join(ROOT, "tests", "cpp", "ref_api_driver")
    assert os.path.exists(exe), "run `make` (or __graft_entry__.build()) first"
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    check_driver1(out.stdout)
