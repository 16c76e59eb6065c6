"""(Named to run last in the GPU suite: it spawns the command lines as subprocesses.)
The command-line drivers on the CUDA shells (spectral_petsc_b200.drivers with its own GpuElliptic / GpuStokes adapters:
device FGMRES, device-assembled preconditioning matrices) against the SAME driver flow over the CPU oracle: identical SNES /
KSP iteration counts (+-1), the same norms of error, the same stokes.vtk numbers."""
import os
import subprocess
import sys

import numpy as np
import pytest

from spectral_petsc_b200 import drivers
from support.oracle_problems import OracleElliptic, OracleStokes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def both(main, cmd, oracle):
    lo, lg = [], []
    ro = main(cmd.split(), out=lo.append, make_problem=oracle)
    rg = main(cmd.split(), out=lg.append)  # default adapters: the GPU
    return ro, rg, lo, lg


def test_elliptic_config1_and_nonlinear(cuda):
    ro, rg, lo, lg = both(drivers.elliptic_main, "-dim 16,16,16 -exact 2 -ksp_rtol 1e-10", OracleElliptic)
    assert lg[:2] == lo[:2]  # header and DOF distribution
    # the reference's in-code default: FGMRES(30) + ILU(2) on the (device-assembled) finite-difference matrix; 16 iterations over the oracle
    assert rg["snes_its"] == ro["snes_its"] == 1 and ro["ksp_its"] == [16] and abs(rg["ksp_its"][0] - 16) <= 1
    assert rg["reason"] == ro["reason"] == "CONVERGED_FNORM_RELATIVE"
    assert rg["exact_residual_abs"] < 5e-11 and rg["error_abs"] < 5e-9
    ro, rg, _, _ = both(drivers.elliptic_main, "-dim 16,16,16 -exact 2 -ksp_rtol 1e-10 -pc_type lu", OracleElliptic)
    assert rg["ksp_its"] == ro["ksp_its"] == [13] and abs(rg["error_abs"] - ro["error_abs"]) < 1e-11 + 1e-3 * ro["error_abs"]
    assert np.abs(rg["x"] - ro["x"]).max() < 1e-9
    # tests.sh: the nonlinear 2-D problem
    ro, rg, _, _ = both(drivers.elliptic_main, "-dim 24,24 -exact 0 -cos_scale 3 -gamma 4 -ksp_rtol 1e-12 -snes_rtol 1e-12 -pc_type lu", OracleElliptic)
    assert abs(rg["snes_its"] - ro["snes_its"]) <= 1 and all(abs(a - b) <= 2 for a, b in zip(rg["ksp_its"], ro["ksp_its"]))
    assert rg["reason"] == ro["reason"] == "CONVERGED_FNORM_RELATIVE" and abs(rg["error_abs"] - ro["error_abs"]) < 1e-9


def test_stokes_continuation_and_vtk(cuda, tmp_path):
    vo, vg = str(tmp_path / "o.vtk"), str(tmp_path / "g.vtk")
    cmd = ("-exact 2 -cont 2 -rheology 1 -eps 1e-2 -exponent 3 -schur_ksp_max_it 3 -vel_ksp_max_it 4 -vel_pc_type hypre -svel_ksp_type preonly "
           "-svel_pc_type hypre -dim 8,8,8 -ksp_rtol 1e-6 -ksp_max_it 300 -output_vtk ")
    lo, lg = [], []
    ro = drivers.stokes_main((cmd + vo).split(), out=lo.append, make_problem=OracleStokes)
    rg = drivers.stokes_main((cmd + vg).split(), out=lg.append)
    assert lg[:3] == lo[:3]
    assert rg["null_space"] < 1e-10
    for a, b in zip(rg["steps"], ro["steps"]):
        assert a["reason"] == b["reason"] == "CONVERGED_FNORM_RELATIVE"
        assert abs(a["snes_its"] - b["snes_its"]) <= 1
        assert all(abs(x - y) <= 1 + y // 10 for x, y in zip(a["ksp_its"], b["ksp_its"]))
        assert abs(a["error"] - b["error"]) < 1e-6 * max(1.0, b["error"])
    # the two files hold the same header lines and the same numbers
    to, tg = open(vo).read().split("\n"), open(vg).read().split("\n")
    assert len(to) == len(tg)
    num = lambda line: [float(t) for t in line.split()]
    for a, b in zip(to, tg):
        if a[:1].isalpha() or a.startswith("#") or not a.strip():
            assert a == b
        else:
            assert np.allclose(num(a), num(b), rtol=1e-5, atol=1e-6)


def test_command_lines(cuda):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "spectral_petsc_b200.elliptic", "-dim", "10,10,10,10", "-pc_type", "hypre", "-exact", "2", "-ksp_monitor",
                        "-ksp_rtol", "1e-10"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)  # README:21 in 4-D (the host LU standing in for hypre would fill 2 GB at 12^5)
    assert r.returncode == 0, r.stderr
    out = r.stdout.split("\n")
    assert out[0] == "Elliptic problem  dims = [10,10,10,10]    gamma = 0.000000    exponent = 2.000000"
    assert out[1] == "DOF distribution:    10000 local         4096 global         5904 dirichlet"
    assert "Number of nonlinear iterations = 1" in out and "Reason for solver termination: CONVERGED_FNORM_RELATIVE" in out
    err = float([l for l in out if l.startswith("Norm of error")][0].split("abs =")[1].split()[0])
    assert err < 1e-8
    r = subprocess.run([sys.executable, "-m", "spectral_petsc_b200.stokes"] + ("-exact 2 -cont0 1 -schur_ksp_max_it 3 -vel_ksp_max_it 4 -vel_pc_type hypre "
                       "-svel_ksp_type preonly -svel_pc_type hypre -ksp_type fgmres -ksp_monitor -dim 20,20,20 -ksp_rtol 1e-10").split(),
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)  # README:44, BASELINE config 4
    assert r.returncode == 0, r.stderr
    out = r.stdout.split("\n")
    assert "DOF distribution: 23328 global   5832/8000 pressure    17496/24000 velocity    6504 dirichlet    0 mixed" in out  # SURVEY 8 header
    assert "Reason for solver termination: CONVERGED_FNORM_RELATIVE" in out
    err = float([l for l in out if l.startswith("Norm of error")][0].split("abs =")[1].split()[0])
    assert err < 1e-6  # 7.8e-08 over the oracle
    r = subprocess.run([sys.executable, "-m", "spectral_petsc_b200.stokes", "-boundary", "2"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 83 and "Boundary type 2 not implemented" in r.stderr


def test_native_elliptic_executable(cuda):
    """apps/elliptic: the reference's ./elliptic in C++ over the reference-API layer, the device FGMRES and the host ILU(2)
    stand-in - no Python in the solve.  Same lines and the same counts as the Python flow over the oracle."""
    exe = os.path.join(ROOT, "apps", "elliptic")
    assert os.path.exists(exe), "run `make` (or __graft_entry__.build()) first"
    cmd = "-dim 16,16,16 -exact 2 -ksp_rtol 1e-10"  # BASELINE.json configs[0]
    r = subprocess.run([exe] + cmd.split(), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    out = r.stdout.split("\n")
    lo = []
    ro = drivers.elliptic_main(cmd.split(), out=lo.append, make_problem=OracleElliptic)
    assert out[:2] == lo[:2]  # header, DOF distribution
    assert out[2].startswith("Norm of exact residual   : abs = ") and float(out[2].split("abs =")[1].split()[0]) < 5e-11
    assert "Number of nonlinear iterations = 1" in out and "Reason for solver termination: CONVERGED_FNORM_RELATIVE" in out
    kits = [int(t) for t in [l for l in out if l.startswith("KSP iterations per Newton step:")][0].split(":")[1].split()]
    assert len(kits) == 1 and abs(kits[0] - ro["ksp_its"][0]) <= 1  # 16 with ILU(2)
    err = float([l for l in out if l.startswith("Norm of error")][0].split("abs =")[1].split()[0])
    assert err < 5e-9
    # the nonlinear problem of tests.sh: same number of Newton steps as the Python flow
    cmd = "-dim 24,24 -exact 0 -cos_scale 3 -gamma 4 -ksp_rtol 1e-10 -snes_rtol 1e-10 -snes_monitor"
    r = subprocess.run([exe] + cmd.split(), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    ro = drivers.elliptic_main(cmd.split(), out=lo.append, make_problem=OracleElliptic)
    its = int([l for l in r.stdout.split("\n") if l.startswith("Number of nonlinear iterations")][0].split("=")[1])
    assert abs(its - ro["snes_its"]) <= 1 and "CONVERGED_FNORM_RELATIVE" in r.stdout
    err = float([l for l in r.stdout.split("\n") if l.startswith("Norm of error")][0].split("abs =")[1].split()[0])
    assert abs(err - ro["error_abs"]) < 1e-8
    # errors follow the reference: unknown exact solution, missing -cos_scale
    assert subprocess.run([exe, "-dim", "8,8", "-exact", "0"], capture_output=True, text=True).returncode == 83


def test_native_stokes_executable(cuda, tmp_path):
    """apps/stokes: the reference's ./stokes in C++ (reference-API layer, device FGMRES for every Krylov solve, host ILU stand-in,
    saddle-point PC / Newton / continuation orchestrated in C++) - same counts, errors and stokes.vtk as the Python flow over the oracle."""
    exe = os.path.join(ROOT, "apps", "stokes")
    assert os.path.exists(exe), "run `make` (or __graft_entry__.build()) first"
    vn, vp = str(tmp_path / "native.vtk"), str(tmp_path / "python.vtk")
    cmd = ("-exact 2 -cont 2 -rheology 1 -eps 1e-2 -exponent 3 -schur_ksp_max_it 3 -vel_ksp_max_it 4 -svel_ksp_type preonly -dim 8,8,8 "
           "-ksp_rtol 1e-6 -ksp_max_it 300 -output_vtk ")
    r = subprocess.run([exe] + (cmd + vn).split(), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr + r.stdout
    out = r.stdout.strip().split("\n")
    lo = []
    ro = drivers.stokes_main((cmd + vp).split(), out=lo.append, make_problem=OracleStokes)
    assert out[:3] == lo[:3]
    assert [l for l in out if l.startswith("## [")] == [l for l in lo if l.startswith("## [")]
    steps = []
    for i, l in enumerate(out):
        if l.startswith("Number of nonlinear iterations"):
            steps.append((int(l.split("=")[1]), out[i + 1].split(": ")[1], float(out[i + 2].split("abs =")[1]), [int(t) for t in out[i + 3].split(":")[1].split()]))
    assert len(steps) == len(ro["steps"]) == 3
    for (its, reason, err, kits), b in zip(steps, ro["steps"]):
        assert reason == b["reason"] == "CONVERGED_FNORM_RELATIVE" and abs(its - b["snes_its"]) <= 1
        assert all(abs(x - y) <= 1 + y // 10 for x, y in zip(kits, b["ksp_its"]))
        assert abs(err - b["error"]) <= 1e-3 * b["error"] + 1e-8
    tn, tp = open(vn).read().split("\n"), open(vp).read().split("\n")
    assert len(tn) == len(tp)
    for a, b in zip(tn, tp):
        if a[:1].isalpha() or a.startswith("#") or not a.strip():
            assert a == b
        else:
            assert np.allclose([float(t) for t in a.split()], [float(t) for t in b.split()], rtol=1e-5, atol=1e-6)
    # the saddle-point PC composed on the device (the default, sb200_saddle_*) against the same composition on host copies
    rh = subprocess.run([exe] + (cmd + vn + " -saddle_on_host 1").split(), capture_output=True, text=True, timeout=900)
    assert rh.returncode == 0, rh.stderr + rh.stdout
    kd = [l for l in out if l.startswith("KSP iterations per Newton step:") or l.startswith("Number of nonlinear iterations")]
    kh = [l for l in rh.stdout.split("\n") if l.startswith("KSP iterations per Newton step:") or l.startswith("Number of nonlinear iterations")]
    assert len(kd) == len(kh) == 6
    for a, b in zip(kd, kh):
        assert all(abs(int(x) - int(y)) <= 1 for x, y in zip(a.replace("=", ":").split(":")[1].split(), b.replace("=", ":").split(":")[1].split()))
    # BASELINE config 4 (README:44) with ILU(2) where the README asks for hypre: 26 outer iterations over the oracle (245 with ILU(0), 11 with LU)
    r = subprocess.run([exe] + ("-exact 2 -cont0 1 -schur_ksp_max_it 3 -vel_ksp_max_it 4 -svel_ksp_type preonly -ksp_type fgmres -dim 20,20,20 -ksp_rtol 1e-10 "
                                "-ksp_max_it 400 -vel_pc_factor_levels 2 -svel_pc_factor_levels 2").split(),
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "DOF distribution: 23328 global   5832/8000 pressure    17496/24000 velocity    6504 dirichlet    0 mixed" in r.stdout
    assert "Reason for solver termination: CONVERGED_FNORM_RELATIVE" in r.stdout
    assert float([l for l in r.stdout.split("\n") if l.startswith("Norm of error")][0].split("abs =")[1]) < 1e-6
    kits = [int(t) for t in [l for l in r.stdout.split("\n") if l.startswith("KSP iterations per Newton step:")][0].split(":")[1].split()]
    assert len(kits) == 1 and abs(kits[0] - 26) <= 2


def test_native_elliptic_tests_sh_sweep(cuda):
    """The reference's tests.sh on the GPU executable: spectral convergence of the nonlinear 2-D problem (errors measured over the CPU
    double: 5.0e-2 / 1.3e-3 / 6.1e-8 / 2.6e-13 at n = 16 / 20 / 28 / 36 for -cos_scale 3)."""
    exe = os.path.join(ROOT, "apps", "elliptic")
    errs = {}
    for n in (16, 20, 28, 36):
        r = subprocess.run([exe] + ("-dim %d,%d -exact 0 -cos_scale 3 -gamma 4 -ksp_rtol 1e-12 -snes_rtol 1e-12" % (n, n)).split(), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr + r.stdout
        errs[n] = float([l for l in r.stdout.split("\n") if l.startswith("Norm of error")][0].split("abs =")[1].split()[0])
    assert abs(errs[16] - 4.9979e-2) < 1e-4 and abs(errs[20] - 1.29596e-3) < 1e-6 and abs(errs[28] - 6.084e-8) < 2e-9 and errs[36] < 1e-11


def test_native_cheb_executable(cuda):
    """apps/cheb: the reference's cheb.c (K1 / K2 of SURVEY 8c) on the GPU through MatCreateChebD1 / MatCreateCheb."""
    exe = os.path.join(ROOT, "apps", "cheb")
    assert os.path.exists(exe), "run `make` (or __graft_entry__.build()) first"

    def norms(args):
        r = subprocess.run([exe] + args.split(), capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr + r.stdout
        return [float(l.split()[-1]) for l in r.stdout.split("\n") if l.startswith("Norm of error")]

    a = norms("")
    assert abs(a[0] - 1.029e-02) < 1e-5 and abs(a[1] - 6.245e-06) < 1e-8
    for axis, want in ((0, 6.245e-06), (1, 8.72e-05), (2, 1.04e-03)):
        assert abs(norms("-m 8 -n 7 -p 6 -d %d" % axis)[1] / want - 1.0) < 2e-3
    assert norms("-m1 16")[0] < 1e-13 and norms("-m 128 -n 128 -p 128 -d 0")[1] < 1e-10
    assert subprocess.run([exe, "-d", "3"], capture_output=True, text=True, timeout=60).returncode == 83
