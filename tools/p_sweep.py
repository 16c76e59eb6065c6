"""Per-P table SURVEY 8d asks for: ChebMult (one axis derivative) and MatMult_Elliptic at P in {16, 32, 48, 64, 96, 128, 129} on
every kernel path that supports the extent, timed with CUDA events (L2 flushed between steps), with both rooflines beside each
number: t_hbm = algorithmic bytes / measured copy bandwidth, t_fp64 = algorithmic flops / measured DMMA peak.  Also times the
device assembly of the finite-difference preconditioning matrix.  One JSON line per row; nothing here reads oracle/.

usage: python tools/p_sweep.py [steps] > gpurun_out/p_sweep.jsonl
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectral_petsc_b200 as sp  # noqa: E402
from tools.time_ops import timeit  # noqa: E402

FP64_TFLOPS = 37.1  # tools/fp64_peak.cu on this pool (profiles/r01_fp64_peak.jsonl)


def hbm_gbs():
    try:
        return float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", 6452.2))
    except Exception:
        return 6452.2


def rows(steps=10, Ps=(16, 32, 48, 64, 96, 128, 129), dev=None, flush=None):
    """Yields one dict per measured row (see the module docstring)."""
    dev = dev or torch.device("cuda:0")
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    bw = hbm_gbs()
    rng = np.random.default_rng(0)
    for P in Ps:
        m = P ** 3
        x = torch.from_numpy(rng.standard_normal(m)).to(dev)
        y = torch.empty_like(x)
        for tr in (0, 2):  # strided and contiguous axis
            C = sp.Cheb(3, tr, [P] * 3)
            ms = timeit(lambda: C.mult(x, y), steps, flush)
            fl, by = 2.0 * P * m, 16.0 * m
            yield {"op": "ChebMult", "P": P, "axis": tr, "ms": ms, "gdof_s": m / ms / 1e6, "t_hbm_ms": by / bw / 1e6, "t_fp64_ms": fl / FP64_TFLOPS / 1e9,
                   "frac_of_binding_roofline": max(by / bw / 1e6, fl / FP64_TFLOPS / 1e9) / ms}
            C.destroy()
        E = sp.Elliptic([P] * 3, gamma=4.0, exponent=2.0)
        us = torch.from_numpy(0.1 * np.random.default_rng(1).standard_normal(E.g)).to(dev)
        E.form_function(us)
        U = torch.from_numpy(rng.standard_normal(E.g)).to(dev)
        V = torch.empty_like(U)
        fl, by = 6 * 2.0 * P * m, 8.0 * (2 * E.g + 5 * m)  # 2d derivatives; U, V, eta, deta, gradu[3] once each
        for path, name in ((1, "generic"), (2, "chain per axis"), (3, "persistent chain")):
            if (path == 2 and P not in (32, 64, 128)) or (path == 3 and not (P % 16 == 0 and 32 <= P <= 160)):
                continue
            E.set_path(path)
            l0 = sp.launch_count()
            E.mat_mult(U, V)
            nl = sp.launch_count() - l0
            ms = timeit(lambda: E.mat_mult(U, V), steps, flush)
            yield {"op": "MatMult_Elliptic", "P": P, "path": name, "launches": nl, "ms": ms, "gdof_s": m / ms / 1e6, "t_hbm_ms": by / bw / 1e6,
                   "t_fp64_ms": fl / FP64_TFLOPS / 1e9, "frac_of_binding_roofline": max(by / bw / 1e6, fl / FP64_TFLOPS / 1e9) / ms}
        E.set_path(0)
        csr = E.jacobian_csr()
        ms_full = timeit(lambda: E.jacobian_csr(), steps, flush)  # includes the torch.empty of the three output arrays
        ms_vals = timeit(lambda: E.jacobian_csr(pattern=csr[:2]), steps, flush)
        nnz = csr[2].numel()
        by_vals = 8.0 * 5 * m + 8.0 * nnz  # eta, deta, gradu[3] once; values written
        yield {"op": "FormJacobian (device CSR)", "P": P, "rows": E.g, "nnz": nnz, "ms_pattern_and_values": ms_full, "ms_values_only": ms_vals,
               "t_hbm_ms_values_only": by_vals / bw / 1e6, "frac_of_hbm_roofline": by_vals / bw / 1e6 / ms_vals}
        E.destroy()


def config_rows(steps=10, dev=None, flush=None):
    """BASELINE configs that are not cubes of the sweep: elliptic 5-D 12^5 (config 3, the README's arbitrary-dimension example) and
    the 16^3 grid of config 1 - both far below the size where a roofline binds, so the launch count is the number to read."""
    dev = dev or torch.device("cuda:0")
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    bw = hbm_gbs()
    for dim in ([12] * 5, [16] * 3):
        d, m = len(dim), int(np.prod(dim))
        E = sp.Elliptic(dim, gamma=4.0, exponent=2.0)
        E.form_function(torch.from_numpy(0.1 * np.random.default_rng(1).standard_normal(E.g)).to(dev))
        U = torch.from_numpy(np.random.default_rng(0).standard_normal(E.g)).to(dev)
        V, F = torch.empty_like(U), torch.empty_like(U)
        fl, by = sum(2 * 2.0 * p for p in dim) * m, 8.0 * (2 * E.g + (2 + d) * m)
        def graphed():  # opt-in path 4: the generic launches replayed from a CUDA graph (measured beside the default)
            E.set_path(4)
            E.mat_mult(U, V)
            E.set_path(0)

        # the opt-in row last: if it raises (first GPU run), the default rows of this grid are already out
        for name, fn in (("MatMult_Elliptic", lambda: E.mat_mult(U, V)), ("FormFunction", lambda: E.form_function(U, F)), ("MatMult_Elliptic (CUDA graph)", graphed)):
            l0 = sp.launch_count()
            fn()
            nl = sp.launch_count() - l0
            ms = timeit(fn, steps, flush)
            yield {"op": name, "dim": "x".join(str(p) for p in dim), "launches": nl, "ms": ms, "gdof_s": m / ms / 1e6, "t_hbm_ms": by / bw / 1e6,
                   "t_fp64_ms": fl / FP64_TFLOPS / 1e9, "frac_of_binding_roofline": max(by / bw / 1e6, fl / FP64_TFLOPS / 1e9) / ms}
        E.destroy()


def config4_rows(steps=10, dev=None, flush=None):
    """BASELINE config 4's grid (stokes 20^3, linear viscosity): the linear shells as launched and as replayed from CUDA graphs
    (sb200_stokes_set_graph, opt-in) - a launch-bound size, the launch count is the number to read."""
    dev = dev or torch.device("cuda:0")
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    dim = [20, 20, 20]
    S = sp.Stokes(dim, rheology=0)
    S.set_dirichlet(torch.zeros(S.dv, dtype=torch.float64, device=dev))
    S.set_force(torch.zeros(S.g, dtype=torch.float64, device=dev))
    rng = np.random.default_rng(0)
    S.function(torch.from_numpy(0.1 * rng.standard_normal(S.g)).to(dev))
    x = torch.from_numpy(rng.standard_normal(S.g)).to(dev)
    xv = torch.from_numpy(rng.standard_normal(S.gv)).to(dev)
    xp = torch.from_numpy(rng.standard_normal(S.gp)).to(dev)
    y, yv, yp = torch.empty_like(x), torch.empty_like(xv), torch.empty_like(xp)
    for name, fn, ndof in (("StokesMatMult", lambda: S.mat_mult(x, y), 4 * S.m), ("StokesMatMultVV", lambda: S.mat_mult_vv(xv, yv), 3 * S.m),
                           ("StokesMatMultPV", lambda: S.mat_mult_pv(xv, yp), 3 * S.m), ("StokesMatMultVP", lambda: S.mat_mult_vp(xp, yv), S.m)):
        for graph in (False, True):
            S.set_graph(graph)
            fn()
            l0 = sp.launch_count()
            fn()
            nl = sp.launch_count() - l0
            ms = timeit(fn, steps, flush)
            yield {"op": name + (" (CUDA graph)" if graph else ""), "dim": "20x20x20", "launches": nl, "ms": ms, "gdof_s": ndof / ms / 1e6}
    S.set_graph(False)
    S.destroy()


def stokes_rows(steps=10, P=128, dev=None, flush=None):
    """The Stokes shells at P^3 (BASELINE config 5 state: -rheology 1 -exponent 3 -eps 1e-4) and the device assembly of MatVVPC."""
    dev = dev or torch.device("cuda:0")
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    bw = hbm_gbs()
    rng = np.random.default_rng(0)
    S = sp.Stokes([P] * 3, rheology=1, hardness=1.0, exponent=3.0, regularization=1e-4, gamma0=1.0)
    xs = torch.from_numpy(0.1 * np.random.default_rng(1).standard_normal(S.g)).to(dev)
    S.set_dirichlet(torch.zeros(S.dv, dtype=torch.float64, device=dev))
    S.set_force(torch.zeros(S.g, dtype=torch.float64, device=dev))
    S.function(xs)
    x = torch.from_numpy(rng.standard_normal(S.g)).to(dev)
    xv = torch.from_numpy(rng.standard_normal(S.gv)).to(dev)
    xp = torch.from_numpy(rng.standard_normal(S.gp)).to(dev)
    y, yv, yp = torch.empty_like(x), torch.empty_like(xv), torch.empty_like(xp)
    m = S.m
    # scalar axis derivatives per application: VV 18, VP 3, PV 3, MatMult / Function 24
    ops = (("StokesMatMult", lambda: S.mat_mult(x, y), 4 * m, 24), ("StokesMatMultVV", lambda: S.mat_mult_vv(xv, yv), 3 * m, 18),
           ("StokesMatMultVP", lambda: S.mat_mult_vp(xp, yv), m, 3), ("StokesMatMultPV", lambda: S.mat_mult_pv(xv, yp), 3 * m, 3),
           ("StokesFunction", lambda: S.function(xs, y), 4 * m, 24))
    for name, fn, ndof, nder in ops:
        l0 = sp.launch_count()
        fn()
        nl = sp.launch_count() - l0
        ms = timeit(fn, steps, flush)
        t_fp64 = nder * 2.0 * P * m / FP64_TFLOPS / 1e9
        yield {"op": name, "P": P, "launches": nl, "ms": ms, "gdof_s": ndof / ms / 1e6, "t_fp64_ms": t_fp64, "frac_of_fp64_roofline": t_fp64 / ms}
    csr = S.pc_velocity_csr()
    ms_vals = timeit(lambda: S.pc_velocity_csr(pattern=csr[:2]), steps, flush)
    nnz = csr[2].numel()
    by_vals = 8.0 * m + 8.0 * nnz  # eta once, values written
    yield {"op": "StokesPCSetUp0 (device CSR)", "P": P, "rows": S.gv, "nnz": nnz, "ms_values_only": ms_vals, "t_hbm_ms_values_only": by_vals / bw / 1e6,
           "frac_of_hbm_roofline": by_vals / bw / 1e6 / ms_vals}
    # the reference's literal sequence of shells (VV, PV, VP each padding and differentiating their input: 24 scalar derivatives)
    # beside the default (pressure rows from the trace of the viscous gradient, pressure folded into the viscous flux: 18)
    S.set_trace_divergence(False)
    S.set_fold_pressure(False)
    for name, fn, ndof, nder in (ops[0], ops[4]):
        l0 = sp.launch_count()
        fn()
        nl = sp.launch_count() - l0
        ms = timeit(fn, steps, flush)
        t_fp64 = nder * 2.0 * P * m / FP64_TFLOPS / 1e9
        yield {"op": name + " (evaluation switches off: three separate shells)", "P": P, "launches": nl, "ms": ms, "gdof_s": ndof / ms / 1e6, "t_fp64_ms": t_fp64,
               "frac_of_fp64_roofline": t_fp64 / ms}
    S.set_fold_pressure(True)
    S.set_trace_divergence(True)
    S.function(xs)  # back to the state the default path leaves
    S.destroy()


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    for row in rows(steps):
        print(json.dumps(row), flush=True)
    for row in stokes_rows(steps):
        print(json.dumps(row), flush=True)
    for row in config_rows(steps):
        print(json.dumps(row), flush=True)
    for row in config4_rows(steps):
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
