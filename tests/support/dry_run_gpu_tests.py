"""Dry run of GPU test modules on CPU: the Python harness (spectral_petsc_b200/capi.py) is pointed at the CPU TEST DOUBLE of the
device entry points (tests/mock/sb200_cpu_double.cpp + the unchanged host layer, built as a shared library) and every "device"
tensor is a CPU tensor.  What this checks is the tests' and the harness's own logic - argument order, sizes, tolerances that do
not depend on the GPU's summation order - before the tests meet a GPU; it says nothing about the CUDA kernels.

usage: python tests/support/dry_run_gpu_tests.py <libmock.so> <module> [test-name-substring ...]
prints one line per test function / parameter set and exits non-zero on the first failure."""
import ctypes
import importlib
import inspect
import itertools
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def patch(libpath):
    import spectral_petsc_b200.capi as capi

    L = ctypes.CDLL(libpath)
    L.sb200_last_error.restype = ctypes.c_char_p
    L.sb200_launch_count.restype = ctypes.c_longlong
    capi._lib = L

    def _ptr(t):
        if not (isinstance(t, torch.Tensor) and t.dtype == torch.float64 and t.is_contiguous()):
            raise TypeError("expected a contiguous fp64 tensor")
        return ctypes.c_void_p(t.data_ptr())

    def _wrap(ptr, n):
        buf = (ctypes.c_double * n).from_address(int(ptr))
        return torch.frombuffer(buf, dtype=torch.float64, count=n)

    capi._ptr = _ptr
    capi._wrap = _wrap
    capi._stream = lambda: None
    # every factory / move that names a CUDA device lands on the CPU instead
    def on_cpu(f):
        def g(*a, **k):
            if str(k.get("device", "")).startswith("cuda"):
                k["device"] = "cpu"
            return f(*a, **k)

        return g

    for name in ("empty", "zeros", "ones", "full", "empty_like", "zeros_like", "ones_like", "full_like", "tensor", "as_tensor", "arange", "randn", "rand"):
        setattr(torch, name, on_cpu(getattr(torch, name)))
    _to = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple(x for x in a if not ((isinstance(x, torch.device) and x.type == "cuda") or (isinstance(x, str) and x.startswith("cuda"))))
        if str(k.get("device", "")).startswith("cuda"):
            k.pop("device")
        return _to(self, *a, **k) if (a or k) else self

    torch.Tensor.to = to
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.is_available = lambda: True
    torch.cuda.current_device = lambda: 0
    torch.cuda.synchronize = lambda *a, **k: None


def param_sets(fn):
    marks = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
    axes = []
    for m in marks:
        names = [n.strip() for n in m.args[0].split(",")] if isinstance(m.args[0], str) else list(m.args[0])
        axes.append([dict(zip(names, v if len(names) > 1 else (v,))) for v in m.args[1]])
    for combo in itertools.product(*axes):
        kw = {}
        for c in combo:
            kw.update(c)
        yield kw


def main():
    libpath, modname, filters = sys.argv[1], sys.argv[2], sys.argv[3:]
    patch(libpath)
    mod = importlib.import_module(modname)
    ran = 0
    for name, fn in inspect.getmembers(mod, inspect.isfunction):
        if not name.startswith("test_") or (filters and not any(f in name for f in filters)):
            continue
        for kw in param_sets(fn):
            args = dict(kw)
            if "cuda" in inspect.signature(fn).parameters:
                args["cuda"] = torch.device("cpu")
            if "tmp_path" in inspect.signature(fn).parameters:
                import pathlib
                import tempfile

                args["tmp_path"] = pathlib.Path(tempfile.mkdtemp(prefix="dryrun"))
            fn(**args)
            ran += 1
            print("ok  %s %s" % (name, kw if kw else ""), flush=True)
    print("dry run: %d test invocations passed" % ran)
    return 0 if ran else 1


if __name__ == "__main__":
    sys.exit(main())
