"""The HOST layer of the product - apps/elliptic.cpp (the native ./elliptic), host/reference_api.cpp, host/petsc_shim.cpp,
host/host_ilu.cpp, csrc/exact.cpp, csrc/fd_rows.h - linked UNCHANGED against tests/mock/sb200_cpu_double.cpp, a CPU test double
of the device-side entry points, and driven end to end: the C++ flow (options, exact solution, FormFunction / FormJacobian
through the reference's names, ILU(2) refresh, FGMRES callbacks, Newton loop, printed lines) gives the same iteration counts
and errors as the Python flow over the oracle.  The double is test infrastructure; GPU parity is tested on the GPU."""
import os
import subprocess

import pytest

from spectral_petsc_b200 import drivers
from support.oracle_problems import OracleElliptic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("native") / "elliptic_cpu_double")
    src = ["apps/elliptic.cpp", "tests/mock/sb200_cpu_double.cpp", "spectral_petsc_b200/host/reference_api.cpp", "spectral_petsc_b200/host/petsc_shim.cpp",
           "spectral_petsc_b200/host/host_ilu.cpp", "spectral_petsc_b200/csrc/exact.cpp", "spectral_petsc_b200/csrc/cheb_matrix.cpp"]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", out] + [os.path.join(ROOT, s) for s in src])
    return out


def native(exe, cmd):
    r = subprocess.run([exe] + cmd.split(), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    out = r.stdout.strip().split("\n")
    val = lambda key: float([l for l in out if l.startswith(key)][0].split("abs =")[1].split()[0])
    kits = [int(t) for t in [l for l in out if l.startswith("KSP iterations per Newton step:")][0].split(":")[1].split()]
    its = int([l for l in out if l.startswith("Number of nonlinear iterations")][0].split("=")[1])
    reason = [l for l in out if l.startswith("Reason for solver termination")][0].split(": ")[1]
    return out, its, kits, reason, val("Norm of error"), val("Norm of exact residual")


CASES = ["-dim 16,16,16 -exact 2 -ksp_rtol 1e-10",                                             # BASELINE configs[0]: FGMRES(30) + ILU(2)
         "-dim 16,16,16 -exact 2 -ksp_rtol 1e-10 -pc_factor_levels 0",
         "-dim 24,24 -exact 0 -cos_scale 3 -gamma 4 -ksp_rtol 1e-12 -snes_rtol 1e-12",           # tests.sh
         "-dim 20,20 -exact 0 -cos_scale 2.8 -gamma 4 -ksp_rtol 1e-12 -snes_rtol 1e-12",
         "-dim 6,6,6,6,6 -exact 2 -ksp_rtol 1e-10 -pc_type jacobi",                              # arbitrary dimension (README:21)
         "-dim 12,12 -exact 1 -pc_type none -ksp_rtol 1e-8 -ksp_gmres_restart 20"]


@pytest.mark.parametrize("cmd", CASES)
def test_native_flow_equals_python_flow(exe, cmd):
    out, its, kits, reason, err, res = native(exe, cmd)
    lines = []
    ro = drivers.elliptic_main(cmd.split(), out=lines.append, make_problem=OracleElliptic)
    assert out[:2] == lines[:2]  # problem header, DOF distribution
    assert (its, reason) == (ro["snes_its"], ro["reason"])
    assert all(abs(a - b) <= 1 for a, b in zip(kits, ro["ksp_its"])) and len(kits) == len(ro["ksp_its"])
    assert abs(err - ro["error_abs"]) <= 1e-6 * ro["error_abs"] + 1e-9  # printed with 7 digits; below 1e-9 the error is the Krylov tolerance
    assert abs(res - ro["exact_residual_abs"]) <= 0.5 * ro["exact_residual_abs"] + 1e-13  # dense-matrix vs FFT derivative: roundoff-level residuals differ


def test_monitors_and_unused_option_warning(exe):
    out, its, kits, _, _, _ = native(exe, "-dim 12,12 -exact 0 -cos_scale 1 -gamma 4 -snes_monitor -ksp_monitor -typo 3")
    assert sum("SNES Function norm" in l for l in out) == its + 1 and sum(l.startswith("    KSP iterations") for l in out) == its
    assert out[-1] == "WARNING! There are options you set that were not used: -typo"
