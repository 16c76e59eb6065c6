"""What the two C++ drivers over the reference's own interface must print (tests/cpp/ref_api_driver*.cpp): shared by the GPU
tests (the real library) and the CPU tests that run the same host layer over the CPU test double."""
import re

import pytest


def check_driver1(txt):
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
    # FormJacobian (elliptic.C:537-590) and StokesPCSetUp0 (stokes.C:1160-1240) through the reference's own names
    fd = re.findall(r"(\w+) P rows (\d+) nz (\d+) sorted (\d) full-stencil rows (\d+)  max \|P x\^2 \+ 2\| = (\S+)  max \|row sum\| = (\S+)", txt)
    assert [(t, int(r), int(z), int(s), int(n)) for t, r, z, s, n, _, _ in fd] == [
        ("elliptic", 2744, 2744 + 3 * 2 * 13 * 14 * 14, 1, 12 ** 3),
        ("elliptic", 100000, 100000 + 5 * 2 * 9 * 10 ** 4, 1, 8 ** 5),
        ("stokes", 3 * 5832, 3 * (5832 + 3 * 2 * 17 * 18 * 18), 1, 3 * 16 ** 3)]
    for row in fd:
        assert float(row[5]) < 1e-8 and float(row[6]) < 1e-8    # 3-point differences are exact for quadratics; entries are O(1e3)
    assert re.findall(r"elliptic P refresh flag (\d) max diff (\S+)", txt) == [("0", "0.000e+00")] * 2


def check_driver2(txt):
    assert "chebD1 vs cheb max diff 0.000e+00" in txt and "chebD1 n=1 -> 83" in txt  # same operator; "n = 1 but must be >= 2" (chebyshev.c:18)
    assert "Schur without an inner solve -> 62" in txt
    m = re.search(r"Schur identity-solve calls (\d+)  max \|S p \+ PV VP p\| / max \|PV VP p\| = (\S+)", txt)
    assert int(m.group(1)) == 1 and float(m.group(2)) < 1e-14
    m = re.search(r"StokesDivergence: \|div\(no bc\) - PV\| = (\S+)  \|div\(exact, with bc\)\| = (\S+)  \|div\(exact, zero bc\)\| = (\S+)", txt)
    assert float(m.group(1)) == 0.0         # withDirichlet = PETSC_FALSE is StokesMatMultPV (stokes.C:557-566)
    assert float(m.group(2)) < 1e-5         # the manufactured velocity is divergence free once its boundary values are in
    assert float(m.group(3)) > 1e-2         # ... and is not when the boundary is zeroed
    # StokesPCApply0..3 over the device-resident composition (host/saddle.cpp), inner solves to 1e-12
    m = re.search(r"null-space removal: pressure mean (\S+)  velocity change (\S+)", txt)
    assert float(m.group(1)) < 1e-15 and float(m.group(2)) == 0.0
    m = re.search(r"saddle type 0 \(block LU, exact inner solves\): max \|PC\(J x\) - x\| / max \|x\| = (\S+)", txt)
    assert float(m.group(1)) < 1e-8         # the block LU factorisation applied exactly inverts StokesMatMult (stokes.C:1712-1713)
    m = re.search(r"saddle type 1 \(upper\): .* = (\S+)", txt)
    assert float(m.group(1)) < 1e-9
    m = re.search(r"saddle type 2 \(diagonal\): max \|VV y_v - r_v\| / max \|r_v\| = (\S+)   \|y_v\(2\) - y_v\(3\)\| = (\S+)   \|y_p\(1\) - y_p\(2\)\| / max = (\S+)", txt)
    assert float(m.group(1)) < 1e-9 and float(m.group(2)) < 1e-12 and float(m.group(3)) < 1e-10
    m = re.search(r"saddle type 3 \(lower\): \|y_p\(3\) - y_p\(0\)\| = (\S+)", txt)
    assert float(m.group(1)) < 1e-10
    m = re.search(r"saddle inner iterations: velocity (\d+)  schur (\d+)", txt)
    assert int(m.group(1)) > 0 and int(m.group(2)) > 0
    assert "saddle wrong-size Vec -> 83   same Vec twice -> 62" in txt
