"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/*.h declares; no compute call is made (no GPU here)."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        txt = open(h).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        syms += re.findall(r"\b(sb200_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(syms))


def test_library_exports_every_declared_symbol():
    import spectral_petsc_b200 as sp

    L = sp.lib()
    syms = declared_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_version_and_error_plumbing():
    import spectral_petsc_b200 as sp

    L = sp.lib()
    assert L.sb200_version() >= 100
    h = ctypes.c_void_p()
    dims = (ctypes.c_int * 2)(4, 4)
    # argument validation happens before any CUDA call and mirrors chebyshev.c:98,106,122
    assert L.sb200_cheb_create(2, 2, dims, ctypes.c_longlong(16), ctypes.byref(h)) == 83
    assert b"tdim out of range" in L.sb200_last_error()
    assert L.sb200_cheb_create(2, 0, dims, ctypes.c_longlong(15), ctypes.byref(h)) == 83
    assert b"dimensions do not agree" in L.sb200_last_error()
    assert L.sb200_cheb_create(1, 0, dims, ctypes.c_longlong(1), ctypes.byref(h)) == 83
    assert b"must be >= 2" in L.sb200_last_error()


def test_no_cpu_fallback_without_gpu():
    import torch

    import spectral_petsc_b200 as sp

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sp.SB200Error) as ei:
        sp.Cheb(1, 0, [8])
    assert ei.value.code == 97


def test_vector_helpers_have_no_cpu_path_either():
    import numpy as np
    import torch

    import spectral_petsc_b200 as sp

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = sp.lib()
    a, b = np.ones(8), np.ones(8)
    pa, pb = a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p)
    assert L.sb200_vec_axpby(ctypes.c_longlong(8), ctypes.c_double(2.0), pa, ctypes.c_double(1.0), pb, None) == 97  # SB200_ERR_CUDA
    assert L.sb200_vec_split(ctypes.c_longlong(2), 3, pa, pb, None, None) == 97
    assert list(b) == [1.0] * 8  # nothing was computed on the host
    assert L.sb200_vec_axpby(ctypes.c_longlong(8), ctypes.c_double(2.0), None, ctypes.c_double(1.0), pb, None) == 62  # argument check first
    assert L.sb200_saddle_create(None, 0, None) == 62 and L.sb200_saddle_apply(None, pa, pb, None) == 62


def test_product_never_imports_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "spectral_petsc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_reference_named_entry_points_exported():
    """The host layer keeps the reference's own function names (sb200_reference_api.h)."""
    import spectral_petsc_b200 as sp

    L = sp.lib()
    txt = open(os.path.join(ROOT, "include", "sb200_reference_api.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"PetscErrorCode\s+([A-Za-z_0-9]+)\s*\(", txt))
    assert {"MatCreateChebD1", "ChebD1Mult", "ChebD1Destroy", "FormJacobian", "StokesPCSetUp0", "MatCreateCheb", "ChebMult", "ChebDestroy", "MatCreate_Elliptic", "MatMult_Elliptic", "FormFunction",
            "StokesCreate", "StokesMatMult", "StokesMatMultVV", "StokesMatMultPV", "StokesMatMultVP", "StokesFunction", "StokesJacobian", "StokesMatMultSchur",
            "StokesMatGetDiagonalSchur", "CreateExactSolution", "StokesCreateExactSolution", "StokesPressureReduceOrder"} <= names
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    shim = open(os.path.join(ROOT, "include", "sb200_petsc_shim.h")).read()
    shim = re.sub(r"/\*.*?\*/", "", shim, flags=re.S)
    for n in set(re.findall(r"PetscErrorCode\s+([A-Za-z_0-9]+)\s*\(", shim)):
        assert hasattr(L, n), n
