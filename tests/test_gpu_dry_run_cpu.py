"""Dry run, on CPU, of the GPU test modules that were written after this round's GPU time was spent: the Python harness is pointed
at the CPU test double of the device entry points (tests/mock/sb200_cpu_double.cpp + the UNCHANGED host layer, as a shared library)
and every "device" tensor is a CPU tensor (tests/support/dry_run_gpu_tests.py).  This exercises the tests' and the harness's own
logic - ctypes signatures, argument order, sizes, the host orchestration in host/saddle.cpp - not the CUDA kernels."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = ["tests/mock/sb200_cpu_double.cpp", "spectral_petsc_b200/host/reference_api.cpp", "spectral_petsc_b200/host/petsc_shim.cpp",
       "spectral_petsc_b200/host/host_ilu.cpp", "spectral_petsc_b200/host/saddle.cpp", "spectral_petsc_b200/csrc/exact.cpp", "spectral_petsc_b200/csrc/cheb_matrix.cpp"]


@pytest.fixture(scope="module")
def libmock(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("mock") / "libsb200_cpu_double.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out] + [os.path.join(ROOT, s) for s in SRC])
    return out


@pytest.mark.parametrize("module,filters,expect", [
    ("test_gpu_zz_saddle", [], 23),
    ("test_gpu_zz_drivers", ["test_elliptic_config1_and_nonlinear", "test_stokes_continuation_and_vtk"], 2),
])
def test_dry_run(libmock, module, filters, expect):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "support", "dry_run_gpu_tests.py"), libmock, module] + filters,
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    assert "dry run: %d test invocations passed" % expect in r.stdout
