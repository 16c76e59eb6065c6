"""Dry run on CPU of what bench.py's child processes execute (tools/p_sweep.py rows, bench.ksp_secondary) over the CPU test double
(see dry_run_gpu_tests.py): their Python logic only - sizes, method names, JSON keys - at small extents; timings are meaningless.

usage: python tests/support/dry_run_bench_extras.py <libmock.so>"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from support.dry_run_gpu_tests import patch  # noqa: E402


class HostEvent:  # stands for torch.cuda.Event
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-6)


def main():
    patch(sys.argv[1])
    torch.cuda.Event = HostEvent
    import bench
    import spectral_petsc_b200 as sp
    import tools.p_sweep as ps

    dev, flush = torch.device("cpu"), torch.empty(1024, dtype=torch.uint8)
    rows = (list(ps.stokes_rows(steps=1, P=16, dev=dev, flush=flush)) + list(ps.config_rows(steps=1, dev=dev, flush=flush)) + list(ps.config4_rows(steps=1, dev=dev, flush=flush))
            + list(ps.rows(steps=1, Ps=(16, 17, 32), dev=dev, flush=flush)))
    G = sp.Elliptic([16, 16, 16], gamma=4.0, exponent=2.0)
    G.form_function(torch.from_numpy(0.1 * np.random.default_rng(1).standard_normal(G.g)))
    ksp = bench.ksp_secondary(sp, torch, dev, G, torch.from_numpy(np.random.default_rng(0).standard_normal(G.g)))
    # the headline flow itself (bench.run_cuda) at 16^3: warm-up, per-step events, the two end-to-end legs, the JSON line
    import argparse
    import contextlib
    import io

    bench.DIM = [16, 16, 16]
    torch.cuda.set_device = lambda *a, **k: None
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.run_cuda(argparse.Namespace(gpus=1, steps=3, warmup=3, path=None, no_cpu_baseline=True, no_extras=True, no_ksp=True, no_stokes=True))
    line = json.loads([l for l in buf.getvalue().splitlines() if l.startswith("{")][-1])
    print(json.dumps({"p_sweep": rows, "ksp": ksp, "line": line}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
