"""Slab-partitioned Stokes shells under torchrun (one rank per GPU): parity against the oracle at a small extent and
timing of StokesMatMult / StokesFunction at 128^3 (BASELINE config 5: -rheology 1 -exponent 3 -eps 1e-4).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 tests/dist/dist_stokes.py 24 128
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))  # repo root
import spectral_petsc_b200 as sp  # noqa: E402
from spectral_petsc_b200 import dist as spd  # noqa: E402


def apply_switches(S):
    """SB200_STOKES_OPTS=trace,fold selects the two evaluation switches (sb200_stokes_set_trace_divergence / _set_fold_pressure;
    both on by default since round 2, SB200_STOKES_OPTS=none turns them off): in slab mode they also remove 2 of the 8 axis-0
    derivative exchanges of a StokesMatMult."""
    opts = os.environ.get("SB200_STOKES_OPTS", "trace,fold").split(",")
    S.set_trace_divergence("trace" in opts)
    S.set_fold_pressure("fold" in opts)
    return [o for o in opts if o in ("trace", "fold")]


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    Pc = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    Pt = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    ok = True
    if Pc:
        from oracle.stokes import StokesCtx  # checker only

        dim = [Pc, Pc, Pc]
        O = StokesCtx(dim, rheology=1, hardness=1.0, exponent=3.0, regularization=1e-2, gamma0=1.0, exact=2)
        O.create_exact_solution()
        S = sp.Stokes(dim, rheology=1, hardness=1.0, exponent=3.0, regularization=1e-2, gamma0=1.0, rank=rank, nranks=world)
        spd.attach_peers(S)
        apply_switches(S)
        S.set_dirichlet(torch.from_numpy(spd.split_dirichlet(O.dirichlet.reshape(-1), dim, world, ncomp=3)[rank].copy()).to(dev))
        S.set_force(torch.from_numpy(spd.split_global(O.force, dim, world, ncomp=4)[rank].copy()).to(dev))
        xs = 0.3 * np.random.default_rng(1).standard_normal(O.g)
        x = np.random.default_rng(0).standard_normal(O.g)
        sl = slice(4 * S.goff, 4 * S.goff + S.g)
        rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
        F = S.function(torch.from_numpy(xs[sl].copy()).to(dev)).cpu().numpy()
        y = S.mat_mult(torch.from_numpy(x[sl].copy()).to(dev)).cpu().numpy()
        e = torch.tensor([rel(F, O.function(xs)[sl]), rel(y, O.mat_mult(x)[sl]), float(S.slab_timeouts())], dtype=torch.float64, device=dev)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        e = e.tolist()
        ok = e[0] < 1e-12 and e[1] < 1e-12 and e[2] == 0
        if rank == 0:
            print(json.dumps({"check": "stokes_slab", "P": Pc, "ranks": world, "function_rel": e[0], "matmult_rel": e[1], "flag_timeouts": e[2], "ok": ok}), flush=True)
        S.destroy()
    if Pt:
        dim = [Pt, Pt, Pt]
        S = sp.Stokes(dim, rheology=1, hardness=1.0, exponent=3.0, regularization=1e-4, gamma0=1.0, rank=rank, nranks=world)
        if world > 1:
            spd.attach_peers(S)
        switches = apply_switches(S)
        S.set_dirichlet(torch.zeros(S.dv, dtype=torch.float64, device=dev))
        S.set_force(torch.zeros(S.g, dtype=torch.float64, device=dev))
        gen = torch.Generator(device=dev).manual_seed(rank)
        xs = 0.1 * torch.randn(S.g, dtype=torch.float64, device=dev, generator=gen)
        x = torch.randn(S.g, dtype=torch.float64, device=dev, generator=gen)
        y = torch.empty_like(x)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        S.function(xs, y)
        res = {}
        for name, fn in (("StokesMatMult", lambda: S.mat_mult(x, y)), ("StokesFunction", lambda: S.function(xs, y))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            ms = []
            for _ in range(10):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
            t = torch.tensor([sorted(ms)[len(ms) // 2]], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[name] = t.item()
        if rank == 0:
            m = Pt ** 3
            print(json.dumps({"bench": "stokes_slab", "P": Pt, "ranks": world, "switches": switches, "ms": res,
                              "gdof_s": {k: 4 * m / v / 1e6 for k, v in res.items()}}), flush=True)
        S.destroy()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
