"""The reference-API host layer (host/reference_api.cpp, host/petsc_shim.cpp, host/host_ilu.cpp, csrc/exact.cpp) linked UNCHANGED
against the CPU test double of the device-side entry points (tests/mock/sb200_cpu_double.cpp) and driven by the same two C++
drivers the GPU tests run: every name of the reference's interface (MatCreateCheb .. StokesPCSetUp0, StokesMatMultSchur) and
its error behaviour are exercised without a GPU, with the same assertions."""
import os
import subprocess

import pytest

from support.ref_api_checks import check_driver1, check_driver2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = ["tests/mock/sb200_cpu_double.cpp", "spectral_petsc_b200/host/reference_api.cpp", "spectral_petsc_b200/host/petsc_shim.cpp",
        "spectral_petsc_b200/host/host_ilu.cpp", "spectral_petsc_b200/host/saddle.cpp", "spectral_petsc_b200/csrc/exact.cpp", "spectral_petsc_b200/csrc/cheb_matrix.cpp"]


@pytest.mark.parametrize("driver,check", [("ref_api_driver", check_driver1), ("ref_api_driver2", check_driver2)])
def test_driver_over_cpu_double(tmp_path, driver, check):
    exe = str(tmp_path / driver)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", driver + ".cpp")] + [os.path.join(ROOT, s) for s in HOST])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr + out.stdout
    check(out.stdout)
