"""Times the device-resident shells (CUDA events, L2 flushed between steps) and prints one JSON line each.

usage: python tools/time_ops.py [stokes|elliptic] [P] [steps]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectral_petsc_b200 as sp  # noqa: E402


def timeit(fn, steps, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2]


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "stokes"
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rng = np.random.default_rng(0)
    if what == "stokes":
        S = sp.Stokes([P] * 3, rheology=1, hardness=1.0, exponent=3.0, regularization=1e-4, gamma0=1.0)
        xs = torch.from_numpy(0.1 * np.random.default_rng(1).standard_normal(S.g)).to(dev)
        S.set_dirichlet(torch.zeros(S.dv, dtype=torch.float64, device=dev))
        S.set_force(torch.zeros(S.g, dtype=torch.float64, device=dev))
        S.function(xs)
        x = torch.from_numpy(rng.standard_normal(S.g)).to(dev)
        xv = torch.from_numpy(rng.standard_normal(S.gv)).to(dev)
        xp = torch.from_numpy(rng.standard_normal(S.gp)).to(dev)
        y, yv, yp = torch.empty_like(x), torch.empty_like(xv), torch.empty_like(xp)
        ops = {
            "StokesMatMult": (lambda: S.mat_mult(x, y), 4 * S.m),
            "StokesMatMultVV": (lambda: S.mat_mult_vv(xv, yv), 3 * S.m),
            "StokesMatMultVP": (lambda: S.mat_mult_vp(xp, yv), S.m),
            "StokesMatMultPV": (lambda: S.mat_mult_pv(xv, yp), 3 * S.m),
            "StokesFunction": (lambda: S.function(xs, y), 4 * S.m),
        }
    else:
        E = sp.Elliptic([P] * 3, gamma=4.0, exponent=2.0)
        us = torch.from_numpy(0.1 * np.random.default_rng(1).standard_normal(E.g)).to(dev)
        F = torch.empty_like(us)
        E.form_function(us, F)
        U = torch.from_numpy(rng.standard_normal(E.g)).to(dev)
        V = torch.empty_like(U)
        ops = {
            "MatMult_Elliptic": (lambda: E.mat_mult(U, V), E.m),
            "FormFunction": (lambda: E.form_function(us, F), E.m),
        }
    for name, (fn, ndof) in ops.items():
        l0 = sp.launch_count()
        fn()
        nl = sp.launch_count() - l0
        ms = timeit(fn, steps, flush)
        print(json.dumps({"op": name, "P": P, "ms": ms, "gdof_s": ndof / ms / 1e6, "launches": nl}))


if __name__ == "__main__":
    main()
