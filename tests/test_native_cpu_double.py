"""The HOST layer of the product - apps/elliptic.cpp (the native ./elliptic), host/reference_api.cpp, host/petsc_shim.cpp,
host/host_ilu.cpp, csrc/exact.cpp, csrc/fd_rows.h - linked UNCHANGED against tests/mock/sb200_cpu_double.cpp, a CPU test double
of the device-side entry points, and driven end to end: the C++ flow (options, exact solution, FormFunction / FormJacobian
through the reference's names, ILU(2) refresh, FGMRES callbacks, Newton loop, printed lines) gives the same iteration counts
and errors as the Python flow over the oracle.  The double is test infrastructure; GPU parity is tested on the GPU."""
import os
import subprocess

import pytest

from spectral_petsc_b200 import drivers
from support.oracle_problems import OracleElliptic, OracleStokes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


HOST = ["tests/mock/sb200_cpu_double.cpp", "spectral_petsc_b200/host/reference_api.cpp", "spectral_petsc_b200/host/petsc_shim.cpp",
        "spectral_petsc_b200/host/host_ilu.cpp", "spectral_petsc_b200/host/saddle.cpp", "spectral_petsc_b200/csrc/exact.cpp", "spectral_petsc_b200/csrc/cheb_matrix.cpp"]


@pytest.fixture(scope="module")
def stokes_exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("native") / "stokes_cpu_double")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", out, os.path.join(ROOT, "apps", "stokes.cpp")] + [os.path.join(ROOT, s) for s in HOST])
    return out


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("native") / "elliptic_cpu_double")
    src = ["apps/elliptic.cpp", "tests/mock/sb200_cpu_double.cpp", "spectral_petsc_b200/host/reference_api.cpp", "spectral_petsc_b200/host/petsc_shim.cpp",
           "spectral_petsc_b200/host/host_ilu.cpp", "spectral_petsc_b200/host/saddle.cpp", "spectral_petsc_b200/csrc/exact.cpp", "spectral_petsc_b200/csrc/cheb_matrix.cpp"]
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


# ---- apps/stokes.cpp ------------------------------------------------------------------------------------------------------------
def native_stokes(exe, cmd):
    r = subprocess.run([exe] + cmd.split(), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr + r.stdout
    out = r.stdout.strip().split("\n")
    steps = []
    for i, l in enumerate(out):
        if l.startswith("Number of nonlinear iterations"):
            steps.append({"snes_its": int(l.split("=")[1]), "reason": out[i + 1].split(": ")[1], "error": float(out[i + 2].split("abs =")[1]),
                          "ksp_its": [int(t) for t in out[i + 3].split(":")[1].split()]})
    return out, steps


BASE = "-schur_ksp_max_it 3 -vel_ksp_max_it 4 -svel_ksp_type preonly"
STOKES_CASES = [
    "-exact 2 -cont0 1 " + BASE + " -ksp_type fgmres -dim 10,10,10 -ksp_rtol 1e-10 -ksp_max_it 200",                                   # README:44 shape, ILU(0) for hypre
    "-exact 2 -cont 2 -rheology 1 -eps 1e-2 -exponent 3 " + BASE + " -dim 8,8,8 -ksp_rtol 1e-6 -ksp_max_it 300",       # README:55 shape (BASELINE config 5)
    "-exact 2 -cont0 1 " + BASE + " -dim 8,8,8 -ksp_rtol 1e-8 -ksp_max_it 200 -pc_saddle_type 1 -vel_pc_factor_levels 2 -svel_pc_factor_levels 2",
    "-exact 2 -cont0 1 " + BASE + " -dim 8,8,8 -ksp_rtol 1e-8 -ksp_max_it 200 -pc_saddle_type 2 -svel_pc_type jacobi",
    "-exact 2 -cont0 1 -schur_ksp_max_it 3 -vel_ksp_max_it 4 -dim 8,8,8 -ksp_rtol 1e-8 -ksp_max_it 200 -pc_saddle_type 3 -vel_pc_factor_levels 1",  # svel: GMRES, not preonly
    "-exact 2 -cont0 1 " + BASE + " -dim 10,10,10 -ksp_rtol 1e-11 -ksp_max_it 400 -vel_pc_type jacobi -svel_pc_type jacobi",  # no host round trip anywhere in the linear solve
]


@pytest.mark.parametrize("cmd", STOKES_CASES)
def test_native_stokes_flow_equals_python_flow(stokes_exe, cmd):
    out, steps = native_stokes(stokes_exe, cmd)
    lines = []
    ro = drivers.stokes_main(cmd.split(), out=lines.append, make_problem=OracleStokes)
    assert out[:3] == lines[:3]  # problem header (2 lines), DOF distribution
    assert [l for l in out if l.startswith("## [")] == [l for l in lines if l.startswith("## [")]  # continuation banners
    assert len(steps) == len(ro["steps"])
    for a, b in zip(steps, ro["steps"]):
        assert (a["snes_its"], a["reason"]) == (b["snes_its"], b["reason"])
        assert len(a["ksp_its"]) == len(b["ksp_its"]) and all(abs(x - y) <= 1 for x, y in zip(a["ksp_its"], b["ksp_its"]))
        assert abs(a["error"] - b["error"]) <= 1e-3 * b["error"] + 1e-8  # the solves stop at their tolerances on both sides


@pytest.mark.parametrize("cmd", STOKES_CASES)
def test_device_resident_saddle_pc_equals_host_orchestrated_one(stokes_exe, cmd):
    """StokesPCApply0..3 composed on device vectors (host/saddle.cpp: sb200_saddle_* over the shells, sb200_ksp and the vector
    helpers - the default) against the same composition written out on host copies in apps/stokes.cpp (-saddle_on_host 1)."""
    _, dev = native_stokes(stokes_exe, cmd)
    _, host = native_stokes(stokes_exe, cmd + " -saddle_on_host 1")
    assert len(dev) == len(host) >= 1
    for a, b in zip(dev, host):
        assert (a["snes_its"], a["reason"], a["ksp_its"]) == (b["snes_its"], b["reason"], b["ksp_its"])
        assert abs(a["error"] - b["error"]) <= 1e-6 * b["error"]


def test_native_stokes_vtk_equals_python_vtk(stokes_exe, tmp_path):
    import numpy as np

    cmd = "-exact 2 -cont 1 -rheology 1 -eps 1e-2 -exponent 2 " + BASE + " -dim 8,8,8 -ksp_rtol 1e-8 -ksp_max_it 200 -snes_max_it 10 -output_vtk "
    vn, vp = str(tmp_path / "native.vtk"), str(tmp_path / "python.vtk")
    native_stokes(stokes_exe, cmd + vn)
    drivers.stokes_main((cmd + vp).split(), out=lambda s: None, make_problem=OracleStokes)
    tn, tp = open(vn).read().split("\n"), open(vp).read().split("\n")
    assert len(tn) == len(tp)
    for a, b in zip(tn, tp):
        if a[:1].isalpha() or a.startswith("#") or not a.strip():
            assert a == b
        else:
            assert np.allclose([float(t) for t in a.split()], [float(t) for t in b.split()], rtol=1e-5, atol=1e-7)


def test_native_stokes_option_errors(stokes_exe):
    for bad in ("-boundary 1", "-rheology 2", "-pcvel 1", "-pc_saddle_type 4", "-ksp_type gmres", "-dim 8,8,8,8", "-exact 4", "-exact 3 -dim 6,6,6", "-vel_pc_type hypre"):
        r = subprocess.run([stokes_exe, "-exact", "2"] + bad.split(), capture_output=True, text=True, timeout=60)
        assert r.returncode == 83 and r.stderr.startswith("error:"), (bad, r.stderr)


def test_native_stokes_exact3_two_dimensional_shear(stokes_exe):
    """-exact 3 (StokesExact3, stokes.C:2016-2034: u = y + 1, v = p = 0, no forcing; 2-D only): the linear field is resolved
    exactly, so one Newton step from zero lands on it to solver tolerance."""
    out, steps = native_stokes(stokes_exe, "-exact 3 -dim 10,8 -cont0 1 " + BASE + " -ksp_rtol 1e-10 -ksp_max_it 200")
    assert "DOF distribution: 144 global   48/80 pressure    96/160 velocity    64 dirichlet    0 mixed" in out
    assert len(steps) == 1 and steps[0]["snes_its"] == 1 and steps[0]["reason"] == "CONVERGED_FNORM_RELATIVE" and steps[0]["error"] < 1e-7


@pytest.mark.parametrize("cos_scale", ["3", "2.8"])
def test_tests_sh_convergence_sweep(exe, cos_scale):
    """The reference's own test script (tests.sh): ./elliptic -dim n,n -exact 0 -cos_scale c -gamma 4 -ksp_rtol 1e-12 -snes_rtol 1e-12 for
    a range of n, reading 'Norm of error'.  The nonlinear problem converges spectrally: the error falls from O(1e-2) at n = 16 to
    rounding level at n = 36..44; the native flow reproduces the Python flow over the oracle."""
    def err(n):
        r = subprocess.run([exe] + ("-dim %d,%d -exact 0 -cos_scale %s -gamma 4 -ksp_rtol 1e-12 -snes_rtol 1e-12" % (n, n, cos_scale)).split(),
                           capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr + r.stdout
        return float([l for l in r.stdout.split("\n") if l.startswith("Norm of error")][0].split("abs =")[1].split()[0])

    e = {n: err(n) for n in (16, 20, 28, 36, 44)}
    assert 1e-3 < e[16] < 1e-1 and e[20] < e[16] / 10 and e[28] < 1e-6 and e[36] < 1e-11 and e[44] < 1e-12
    lines = []
    ro = drivers.elliptic_main(("-dim 20,20 -exact 0 -cos_scale %s -gamma 4 -ksp_rtol 1e-12 -snes_rtol 1e-12" % cos_scale).split(), out=lines.append,
                               make_problem=OracleElliptic)
    assert abs(e[20] - ro["error_abs"]) <= 1e-6 * ro["error_abs"] + 1e-12


# ---- apps/cheb.cpp: the reference's cheb.c ---------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cheb_exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("native") / "cheb_cpu_double")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", out, os.path.join(ROOT, "apps", "cheb.cpp")] + [os.path.join(ROOT, s) for s in HOST])
    return out


def cheb_norms(exe, args):
    r = subprocess.run([exe] + args.split(), capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr + r.stdout
    return [float(l.split()[-1]) for l in r.stdout.split("\n") if l.startswith("Norm of error")]


def test_native_cheb_known_answers(cheb_exe):
    """cheb.c's two printed norms: 1.029e-02 at the default m1 = 5 with spectral decay, and 6.245e-06 / 8.72e-05 /
    1.04e-03 for the derivative of exp(x) + exp(y) + exp(z) along axes 0 / 1 / 2 of the (8, 7, 6) grid."""
    assert cheb_norms(cheb_exe, "") == pytest.approx([1.029e-02, 6.245e-06], rel=1e-3)  # defaults: m1 5, (8, 7, 1), axis 0
    assert cheb_norms(cheb_exe, "-m1 8")[0] == pytest.approx(6.2e-06, rel=2e-2) and cheb_norms(cheb_exe, "-m1 16")[0] < 1e-13
    for axis, want in ((0, 6.245e-06), (1, 8.72e-05), (2, 1.04e-03)):
        assert cheb_norms(cheb_exe, "-m 8 -n 7 -p 6 -d %d" % axis)[1] == pytest.approx(want, rel=2e-3)
    assert cheb_norms(cheb_exe, "-m 24 -n 20 -p 18 -d 2")[1] < 1e-12
    r = subprocess.run([cheb_exe, "-d", "2"], capture_output=True, text=True, timeout=60)  # p = 1 by default: nothing to differentiate along axis 2
    assert r.returncode == 83 and "must be >= 2" in r.stderr
    r = subprocess.run([cheb_exe, "-d", "3"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 83 and "tdim out of range" in r.stderr  # chebyshev.c:106


def test_host_layer_is_clean_under_address_and_ub_sanitizers(tmp_path):
    """apps/stokes + the unchanged host layer (reference_api, saddle, petsc shim, ILU) over the test double, built with
    -fsanitize=address,undefined: a nonlinear continuation with the device-composed block-LU PC, a full GMRES as KSPSchurVelocity
    and the VTK dump must run without a report (heap errors, leaks, undefined behaviour)."""
    exe = str(tmp_path / "stokes_asan")
    cc = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-o", exe,
                         os.path.join(ROOT, "apps", "stokes.cpp")] + [os.path.join(ROOT, s) for s in HOST], capture_output=True, text=True)
    if cc.returncode != 0:
        pytest.skip("sanitizer runtime not available: " + cc.stderr[-200:])
    cmd = ("-exact 2 -cont 1 -rheology 1 -eps 1e-2 -exponent 2 -schur_ksp_max_it 3 -vel_ksp_max_it 4 -dim 7,7,7 -ksp_rtol 1e-8 -ksp_max_it 200 "
           "-snes_max_it 10 -output_vtk " + str(tmp_path / "a.vtk"))
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:halt_on_error=1", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
    r = subprocess.run([exe] + cmd.split(), capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr and "LeakSanitizer" not in r.stderr
    assert r.stdout.count("Reason for solver termination: CONVERGED_FNORM_RELATIVE") == 2
