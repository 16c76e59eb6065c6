"""Host-only entry points of the reference-API layer (include/sb200_reference_api.h) that need no device: util.C's polyInterp with
its own self test (util.C:146-171), the rheology callbacks StokesOptions stores (stokes.C:1920-1944), StokesJacobian's
flag; and that the library exports every reference-API name the header declares."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

import spectral_petsc_b200 as sp
from oracle.stokes import poly_interp, rheology_power

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = ctypes.c_double
PD = ctypes.POINTER(ctypes.c_double)


class StokesOptionsB200(ctypes.Structure):  # include/sb200_reference_api.h (stokes.C:67-79)
    _fields_ = [("numDims", ctypes.c_int), ("dim", ctypes.c_int * 3), ("exact", ctypes.c_int), ("rheology", ctypes.c_int),
                ("hardness", D), ("exponent", D), ("regularization", D), ("gamma0", D)]


def _poly(L, n, x, f, x0, x1):
    w = np.zeros(4 * n)
    w[0::4] = f
    w[1::4] = f
    f0, f1 = D(), D()
    L.polyInterp.argtypes = [ctypes.c_int, PD, PD, D, D, PD, PD]
    rc = L.polyInterp(n, np.ascontiguousarray(x, dtype=np.float64).ctypes.data_as(PD), w.ctypes.data_as(PD), x0, x1, ctypes.byref(f0), ctypes.byref(f1))
    assert rc == 0
    return f0.value, f1.value


def test_reference_api_header_names_are_exported():
    txt = open(os.path.join(ROOT, "include", "sb200_reference_api.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = sorted(set(re.findall(r"^PetscErrorCode\s+([A-Za-z_0-9]+)\s*\(", txt, flags=re.M)))
    assert len(names) >= 30 and "MatMult_Elliptic" in names and "StokesMatMult" in names and "polyInterp" in names
    L = sp.lib()
    assert not [s for s in names if not hasattr(L, s)]


@pytest.mark.parametrize("order", range(2, 20))
def test_polyinterp_self_test_of_util_C(order):
    """util.C:155-171: cos on the nodes 1..order, evaluated at 1.43 and 3.1."""
    x = 1.0 + np.arange(order)
    f0, f1 = _poly(sp.lib(), order, x, np.cos(x), 1.43, 3.1)
    o0, o1 = poly_interp(order, x, np.cos(x), 1.43, 3.1)
    assert f0 == pytest.approx(float(o0), rel=1e-14, abs=1e-15) and f1 == pytest.approx(float(o1), rel=1e-14, abs=1e-15)
    for xe, fe in ((1.43, f0), (3.1, f1)):  # the interpolation error bound prod|x - x_i| / n! * max|cos^(n)|
        assert abs(fe - math.cos(xe)) <= np.prod(np.abs(xe - x)) / math.factorial(order) + 1e-12
    if order >= 3:  # 3.1 lies next to the node 3: already close at low order, and exact at a node
        assert _poly(sp.lib(), order, x, np.cos(x), 2.0, 3.0) == (pytest.approx(math.cos(2.0), abs=1e-14), pytest.approx(math.cos(3.0), abs=1e-14))


def test_polyinterp_reproduces_polynomials_and_pressure_extrapolation_nodes():
    # the use in StokesPressureReduceOrder (stokes.C:1042-1074): interior Chebyshev nodes, evaluated at both ends +-1
    for m in (4, 7, 12, 33):
        xi = np.cos(np.arange(1, m - 1) * np.pi / (m - 1))
        c = np.random.default_rng(m).standard_normal(m - 2)
        f = np.polyval(c, xi)  # degree m-3: reproduced exactly by m-2 nodes
        f0, f1 = _poly(sp.lib(), m - 2, xi, f, 1.0, -1.0)
        scale = np.abs(c).sum()
        assert abs(f0 - np.polyval(c, 1.0)) < 1e-9 * scale * m and abs(f1 - np.polyval(c, -1.0)) < 1e-9 * scale * m
    f0, f1 = _poly(sp.lib(), 1, [0.3], [2.5], 1.0, -1.0)  # one node: the constant
    assert (f0, f1) == (2.5, 2.5)
    assert sp.lib().polyInterp(0, None, None, D(0), D(0), None, None) != 0


@pytest.mark.parametrize("exponent,reg,hard,g0", [(3.0, 1e-4, 1.0, 1.0), (1.0, 1.0, 1.0, 1.0), (2.2, 1e-2, 3.5, 0.7), (1e-6, 0.5, 1.0, 1.0)])
def test_rheology_callbacks(exponent, reg, hard, g0):
    L = sp.lib()
    opt = StokesOptionsB200(3, (8, 8, 8), 0, 1, hard, exponent, reg, g0)
    eta, deta = D(), D()
    L.StokesRheologyPower.argtypes = [ctypes.c_int, D, PD, PD, ctypes.c_void_p]
    L.StokesRheologyLinear.argtypes = [ctypes.c_int, D, PD, PD, ctypes.c_void_p]
    for gamma in (0.0, 1e-8, 0.3, 2.0, 150.0):
        assert L.StokesRheologyPower(3, gamma, ctypes.byref(eta), ctypes.byref(deta), ctypes.addressof(opt)) == 0
        with np.errstate(all="ignore"):
            e, de = rheology_power(np.float64(gamma), hard, exponent, reg, g0)
        assert eta.value == pytest.approx(float(e), rel=1e-15) and deta.value == pytest.approx(float(de), rel=1e-15, abs=0.0)
        assert L.StokesRheologyLinear(3, gamma, ctypes.byref(eta), ctypes.byref(deta), None) == 0
        assert (eta.value, deta.value) == (1.0, 0.0)
    if exponent == 1.0:  # Newtonian limit of the power law
        assert (eta.value, deta.value) == (1.0, 0.0)


def test_stokes_jacobian_reports_different_nonzero_pattern():
    flag = ctypes.c_int(-1)
    assert sp.lib().StokesJacobian(None, None, None, None, ctypes.byref(flag), None) == 0  # stokes.C:761-769
    txt = open(os.path.join(ROOT, "include", "sb200_petsc_shim.h")).read()
    m = re.search(r"DIFFERENT_NONZERO_PATTERN\s*=\s*(\d+)", txt)
    assert m is None or flag.value == int(m.group(1))
    assert flag.value != -1


class StokesExactBoundaryCtx(ctypes.Structure):
    _fields_ = [("exact", ctypes.c_void_p), ("exactCtx", ctypes.c_void_p)]


@pytest.mark.parametrize("exact,d", [(0, 2), (0, 3), (1, 2), (1, 3), (2, 2), (2, 3), (3, 2)])
def test_stokes_exact_point_functions(exact, d):
    """StokesExact0..3 / StokesDirichlet (stokes.C:1948-2050) as per-point host callbacks: the numbers StokesCreateExactSolution puts
    into the vectors (sb200_stokes_exact_solution, compared with the oracle in tests/test_exact_solutions_cpu.py)."""
    L = sp.lib()
    dim = [6, 5, 4][:d]
    U, U2, dirichlet = sp.stokes_exact_solution(dim, exact)
    fn = getattr(L, "StokesExact%d" % exact)
    fn.argtypes = [ctypes.c_int, PD, PD, PD, ctypes.c_void_p]
    nodes = np.stack(np.meshgrid(*[np.cos(np.arange(p) * np.pi / (p - 1)) for p in dim], indexing="ij"), axis=-1).reshape(-1, d)
    bdy = np.zeros(len(nodes), dtype=bool)
    for j, p in enumerate(dim):
        idx = np.indices(dim)[j].reshape(-1)
        bdy |= (idx == 0) | (idx == p - 1)
    val, rhs, bval = (ctypes.c_double * 4)(), (ctypes.c_double * 4)(), (ctypes.c_double * 4)()
    bctx = StokesExactBoundaryCtx(ctypes.cast(fn, ctypes.c_void_p), None)
    L.StokesDirichlet.argtypes = [ctypes.c_int, PD, PD, ctypes.POINTER(ctypes.c_int), PD, ctypes.c_void_p]
    gi = di = 0
    for c, b in zip(nodes, bdy):
        cc = (ctypes.c_double * 3)(*c)
        assert fn(d, cc, val, rhs, None) == 0
        if b:
            kind = ctypes.c_int(-1)
            assert L.StokesDirichlet(d, cc, None, ctypes.byref(kind), bval, ctypes.addressof(bctx)) == 0 and kind.value == 0  # DIRICHLET
            assert list(bval[:d]) == list(dirichlet.reshape(-1, d)[di])
            di += 1
        else:
            assert list(val[:d + 1]) == list(U.reshape(-1, d + 1)[gi]) and list(rhs[:d + 1]) == list(U2.reshape(-1, d + 1)[gi])
            gi += 1
    assert gi * (d + 1) == U.size and di * d == dirichlet.size
    if exact == 3:
        assert L.StokesExact3(3, (ctypes.c_double * 3)(0.1, 0.2, 0.3), val, rhs, None) != 0  # "only implemented for dimension 2" (stokes.C:2021)
