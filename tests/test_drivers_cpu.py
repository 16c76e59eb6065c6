"""The command-line drivers (spectral_petsc_b200/drivers.py = main() of elliptic.C and stokes.C): option handling, the Newton /
continuation flow, the printed lines and stokes.vtk, run here over the CPU oracle through tests/support/oracle_problems.py
(the flows themselves are backend-agnostic; tests/test_gpu_drivers.py runs them on the CUDA shells and compares)."""
import re

import numpy as np
import pytest

from spectral_petsc_b200 import drivers
from support.oracle_problems import OracleElliptic, OracleStokes


def run(main, cmd, make):
    lines = []
    res = main(cmd.split(), out=lines.append, make_problem=make)
    return res, lines


def test_options_database_slice():
    o = drivers.PetscOptions("-dim 12,12,12 -exact 2 -ksp_monitor -ksp_rtol 1e-10 -shift -1.5 -snes_monitor".split())
    assert o.int_array("dim", [8, 6]) == [12, 12, 12] and o.int("exact", 0) == 2 and o.real("ksp_rtol", 1e-5) == 1e-10
    assert o.has("ksp_monitor") and o.has("snes_monitor") and not o.has("output_vtk")
    assert o.real("shift", 0.0) == -1.5  # a negative value is a value, not the next option
    assert o.int_array("other", [8, 6]) == [8, 6] and o.real("gamma", 0.25) == 0.25
    assert drivers.PetscOptions(["-a", "1", "-b"]).unused() == ["a", "b"]
    with pytest.raises(drivers.OptionsError):
        drivers.PetscOptions(["dim", "3"])
    with pytest.raises(drivers.OptionsError):
        drivers.PetscOptions(["-dim", "1,2,3,4,5,6,7,8,9,10,11"]).int_array("dim", [8, 6])  # elliptic.C:138: at most 10


def test_elliptic_config1_lines_and_counts():
    """BASELINE config 1: ./elliptic -dim 16,16,16 -exact 2 -ksp_rtol 1e-10."""
    res, lines = run(drivers.elliptic_main, "-dim 16,16,16 -exact 2 -ksp_rtol 1e-10", OracleElliptic)
    assert lines[0] == "Elliptic problem  dims = [16,16,16]    gamma = 0.000000    exponent = 2.000000"
    assert lines[1] == "DOF distribution:     4096 local         2744 global         1352 dirichlet"  # elliptic.C:424
    assert re.match(r"Norm of exact residual   : abs = \S+   rel = \S+$", lines[2]) and res["exact_residual_abs"] < 5e-11  # K3
    assert lines[-3:-1] == ["Number of nonlinear iterations = 1", "Reason for solver termination: CONVERGED_FNORM_RELATIVE"]
    assert re.match(r"Norm of error            : abs = \S+   rel = \S+$", lines[-1])
    # the reference's in-code solver: FGMRES(30) preconditioned by ILU(2) of the finite-difference matrix (elliptic.C:181-184)
    assert res["ksp_its"] == [16] and res["error_abs"] < 2e-9
    res_lu, _ = run(drivers.elliptic_main, "-dim 16,16,16 -exact 2 -ksp_rtol 1e-10 -pc_type lu", OracleElliptic)
    res_0, _ = run(drivers.elliptic_main, "-dim 16,16,16 -exact 2 -ksp_rtol 1e-10 -pc_factor_levels 0", OracleElliptic)
    assert res_lu["ksp_its"] == [13] and res_0["ksp_its"] == [25]  # a stronger / weaker PC on the same matrix


def test_elliptic_tests_sh_case_and_5d():
    # tests.sh: ./elliptic -dim n,n -exact 0 -cos_scale {3,2.8} -gamma 4 -ksp_rtol 1e-12 -snes_rtol 1e-12
    errs = {}
    for cs, n in ((3, 16), (3, 24), (3, 32), (2.8, 32)):
        res, lines = run(drivers.elliptic_main, "-dim %d,%d -exact 0 -cos_scale %s -gamma 4 -ksp_rtol 1e-12 -snes_rtol 1e-12 -snes_monitor" % (n, n, cs),
                         OracleElliptic)
        assert res["reason"] == "CONVERGED_FNORM_RELATIVE" and 2 <= res["snes_its"] <= 12
        assert sum(l.startswith("  ") and "SNES Function norm" in l for l in lines) == res["snes_its"] + 1
        errs[(cs, n)] = res["error_abs"]
    # what the sweep of tests.sh shows: spectral convergence of the NONLINEAR problem in n (5e-2, 1e-5, 2e-10 for cos_scale 3)
    assert errs[(3, 16)] > 1e3 * errs[(3, 24)] > 1e6 * errs[(3, 32)] and errs[(3, 32)] < 1e-9 and errs[(2.8, 32)] < 1e-8
    # README:21 at a smaller extent: arbitrary dimension
    res, _ = run(drivers.elliptic_main, "-dim 6,6,6,6,6 -exact 2 -ksp_rtol 1e-10 -pc_type lu", OracleElliptic)
    assert res["g"] == 4 ** 5 and res["snes_its"] == 1
    with pytest.raises(drivers.OptionsError):
        run(drivers.elliptic_main, "-dim 8,8 -exact 0", OracleElliptic)  # -cos_scale has no default upstream
    with pytest.raises(drivers.OptionsError):
        run(drivers.elliptic_main, "-dim 8,8 -exact 1 -pc_type sor", OracleElliptic)
    res, lines = run(drivers.elliptic_main, "-dim 8,8 -exact 1 -pc_type jacobi -ksp_rtol 1e-8 -typo 3", OracleElliptic)
    assert lines[-1] == "WARNING! There are options you set that were not used: -typo"


README_STOKES = "-exact 2 -cont0 1 -schur_ksp_max_it 3 -vel_ksp_max_it 4 -vel_pc_type hypre -svel_ksp_type preonly -svel_pc_type hypre -ksp_type fgmres -ksp_rtol 1e-10"


def test_stokes_readme_line_small():
    res, lines = run(drivers.stokes_main, README_STOKES + " -dim 10,10,10", OracleStokes)
    assert lines[0] == "Stokes problem  dim = [10,10,10]"
    assert lines[1] == "  hardness = 1.000000    exponent = 1.000000    regularization = 1.000000    gamma0 = 1.000000"
    assert lines[2] == "DOF distribution: 2048 global   512/1000 pressure    1536/3000 velocity    1464 dirichlet    0 mixed"  # stokes.C:891
    assert any(l.startswith("Norm of solution") for l in lines) and res["exact_residual"] < 1e-5 and res["null_space"] < 1e-10  # 10 nodes per axis: truncation ~3e-6 (2e-11 at 20^3)
    assert "## [1/1] Solving with exponent = 1.000000 regularization 1.00e+00" in lines  # -cont0 1: one solve, the final parameters
    assert len(res["steps"]) == 1 and res["steps"][0]["reason"] == "CONVERGED_FNORM_RELATIVE" and res["steps"][0]["snes_its"] == 1
    assert res["steps"][0]["error"] < 1e-5
    assert lines[-1].startswith("Norm of error            : abs = ")


def test_stokes_continuation_and_vtk(tmp_path):
    """README:55 (BASELINE config 5) at a small extent, with -output_vtk."""
    vtk = str(tmp_path / "stokes.vtk")
    cmd = ("-exact 2 -cont 2 -rheology 1 -eps 1e-2 -exponent 3 -schur_ksp_max_it 3 -vel_ksp_max_it 4 -vel_pc_type hypre -svel_ksp_type preonly "
           "-svel_pc_type hypre -dim 8,8,8 -ksp_rtol 1e-6 -ksp_max_it 300 -output_vtk " + vtk)
    res, lines = run(drivers.stokes_main, cmd, OracleStokes)
    assert [s["step"] for s in res["steps"]] == [0, 1, 2]
    assert res["steps"][0]["exponent"] == 1.0 and res["steps"][2]["exponent"] == 3.0 and res["steps"][2]["regularization"] == pytest.approx(1e-2)
    assert "## [1/2] Solving with exponent = 2.148698 regularization 1.00e-01" in lines  # stokes.C:218-220
    assert all(s["reason"] == "CONVERGED_FNORM_RELATIVE" for s in res["steps"])
    assert [s["snes_its"] for s in res["steps"]] == [2, 3, 4] and all(k <= 12 for s in res["steps"] for k in s["ksp_its"])
    assert res["steps"][0]["error"] < 1e-3  # step 0 is the linear problem the exact solution belongs to (README:48-50)
    assert sum(l.startswith("Minimum eta = ") for l in lines) >= 10  # every residual evaluation prints it (stokes.C:731-734)
    txt = open(vtk).read().split("\n")
    assert txt[:6] == ["# vtk DataFile Version 2.0", "Stokes Output", "ASCII", "DATASET STRUCTURED_GRID", "DIMENSIONS 8 8 8", "POINTS 512 double"]
    assert [float(t) for t in txt[6].split()] == [1.0, 1.0, 1.0]  # node 0 is xi = +1 on every axis
    for head in ("POINT_DATA 512", "VECTORS velocity double", "SCALARS pressure double 1", "VECTORS vel_force double", "SCALARS div_force double 1",
                 "SCALARS eta double 1", "SCALARS deta double 1", "TENSORS strain double"):
        assert head in txt
    i = txt.index("TENSORS strain double")
    assert len(txt[i + 1:]) == 512 * 4 + 1 and len(txt[i + 1].split()) == 3
    # the velocity block carries the Dirichlet values on the boundary: node 0 = the exact solution at (1, 1, 1)
    j = txt.index("VECTORS velocity double")
    u0 = [float(t) for t in txt[j + 1].split()]
    assert u0[0] == pytest.approx(np.sin(0.5 * np.pi) * np.cos(0.5 * np.pi), abs=1e-12) and u0[1] == pytest.approx(-np.cos(0.5 * np.pi), abs=1e-12) and u0[2] == 0
    # eta of the final (power-law) state is not constant
    k = txt.index("SCALARS eta double 1")
    eta = np.array([float(t) for t in txt[k + 2:k + 2 + 512]])
    assert eta.min() > 0 and eta.max() > 1.5 * eta.min()


def test_vtk_writer_two_dimensional_padding(tmp_path):
    # StokesVecView pads 2-D vectors to 3 columns with the literal "0 " (stokes.C:1907-1909); tensors are 3 x 3 with zeros
    dim, nodes, d = [4, 3], 12, 2
    rng = np.random.default_rng(0)
    path = str(tmp_path / "s.vtk")
    strain = [rng.standard_normal((nodes, d)) for _ in range(d)]
    drivers.write_stokes_vtk(path, dim, rng.standard_normal((nodes, d)), rng.standard_normal((nodes, d)), rng.standard_normal(nodes),
                             np.zeros((nodes, d)), np.zeros(nodes), np.ones(nodes), np.zeros(nodes), strain)
    txt = open(path).read().split("\n")
    assert txt[4] == "DIMENSIONS 4 3 1" and txt[6].endswith(" 0 ") and len(txt[6].split()) == 3
    i = txt.index("TENSORS strain double")
    row0, row2 = txt[i + 1].split(), txt[i + 3].split()
    assert float(row0[0]) == pytest.approx(strain[0][0, 0], rel=1e-6) and float(row0[2]) == 0.0 and [float(t) for t in row2] == [0.0, 0.0, 0.0]
    assert txt[txt.index("SCALARS eta double 1") + 2].strip() == "%e" % 1.0


def test_stokes_option_errors():
    for bad in ("-boundary 1", "-rheology 2", "-pcvel 1", "-pc_saddle_type 4", "-ksp_type gmres", "-dim 8,8,8,8"):
        with pytest.raises(drivers.OptionsError):
            run(drivers.stokes_main, "-exact 2 " + bad, OracleStokes)


def test_command_line_refuses_to_run_without_a_gpu():
    """No CPU fallback: on a box without a CUDA device the command line fails loudly instead of computing on the host."""
    import os
    import subprocess
    import sys

    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "spectral_petsc_b200.elliptic", "-dim", "8,8", "-exact", "1"], capture_output=True, text=True, timeout=300,
                       cwd=root, env=dict(os.environ, PYTHONPATH=root))
    assert r.returncode != 0 and "no CUDA device" in r.stderr
    r = subprocess.run([sys.executable, "-m", "spectral_petsc_b200.stokes", "-pcvel", "2"], capture_output=True, text=True, timeout=300,
                       cwd=root, env=dict(os.environ, PYTHONPATH=root))
    assert r.returncode == 83 and "pcvel type number 2 not implemented" in r.stderr


def test_native_executable_option_errors_and_no_fallback():
    """apps/elliptic (C++): option errors are PETSC_ERR_USER before any device work; without a GPU the run fails loudly."""
    import os
    import subprocess

    import torch

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "apps", "elliptic")
    assert os.path.exists(exe), "run `make` first"
    for bad in (["-dim", "8,8", "-exact", "0"], ["dim"], ["-dim", "1,2,3,4,5,6,7,8,9,10,11"], ["-pc_type", "sor", "-exact", "1"], ["-ksp_type", "cg", "-exact", "1"]):
        r = subprocess.run([exe] + bad, capture_output=True, text=True, timeout=60)
        assert r.returncode == 83 and r.stderr.startswith("error:"), (bad, r.stderr)
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "-dim", "8,8", "-exact", "1"], capture_output=True, text=True, timeout=60)
        assert r.returncode == 97 and "error 97" in r.stderr  # SB200_ERR_CUDA: no device, no CPU path
        assert r.stdout.startswith("Elliptic problem  dims = [8,8]    gamma = 0.000000    exponent = 2.000000")
