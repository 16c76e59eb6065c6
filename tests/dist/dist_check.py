"""Multi-process parity check of the slab partition: run under torchrun, one rank per GPU.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tests/dist/dist_check.py 32 64 128

Every rank builds its slab context, the CUDA IPC handles are exchanged through torch.distributed, and
FormFunction / MatMult_Elliptic (generic and fused slab paths) are compared with the oracle's
single-domain result (bar 1e-12 max-norm relative).  Prints one JSON line per extent on rank 0.
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
from oracle.elliptic import MatElliptic  # noqa: E402  (checker only)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for P in [int(a) for a in sys.argv[1:]] or [32]:
        dim = [P, P, P]
        O = MatElliptic(dim, gamma=4.0, exponent=2.0, workers=4)
        O.create_exact_solution(2)
        C = sp.Elliptic(dim, gamma=4.0, exponent=2.0, rank=rank, nranks=world)
        spd.attach_peers(C)
        C.set_dirichlet(torch.from_numpy(spd.split_dirichlet(O.dirichlet, dim, world)[rank].copy()).to(dev))
        C.set_rhs(torch.from_numpy(spd.split_global(O.b, dim, world)[rank].copy()).to(dev))
        Us = 0.1 * np.random.default_rng(1).standard_normal(O.g)
        U = np.random.default_rng(0).standard_normal(O.g)
        sl = slice(C.goff, C.goff + C.g)
        rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
        Fo = O.form_function(Us)
        F = C.form_function(torch.from_numpy(Us[sl].copy()).to(dev)).cpu().numpy()
        Vo = O.mat_mult(U)
        Ud = torch.from_numpy(U[sl].copy()).to(dev)
        V = C.mat_mult(Ud).cpu().numpy()
        V2 = C.mat_mult(Ud).cpu().numpy()
        C.set_path(1)
        Vg = C.mat_mult(Ud).cpu().numpy()
        errs = torch.tensor([rel(F, Fo[sl]), rel(V, Vo[sl]), rel(Vg, Vo[sl]), float(not np.array_equal(V, V2)), float(C.slab_timeouts())],
                            dtype=torch.float64, device=dev)
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        e = errs.tolist()
        good = e[0] < 1e-12 and e[1] < 1e-12 and e[2] < 1e-12 and e[3] == 0 and e[4] == 0
        ok = ok and good
        if rank == 0:
            print(json.dumps({"check": "slab", "P": P, "ranks": world, "function_rel": e[0], "matmult_fused_rel": e[1],
                              "matmult_generic_rel": e[2], "repeat_differs": e[3], "flag_timeouts": e[4], "ok": good}), flush=True)
        if P <= 32:
            # slab-partitioned FGMRES: dot products through the peer-memory all-reduce; counts equal the single-domain oracle solve
            from oracle.fgmres import fgmres  # checker only

            b = np.random.default_rng(3).standard_normal(O.g)
            xo, its_o, hist_o, reason_o = fgmres(O.mat_mult, b, rtol=1e-6, maxits=60)
            C.set_path(0)
            K = sp.KSP(C.g, rank=rank, nranks=world)
            spd.attach_peers(K)
            K.set_operators(C)
            K.set_tolerances(rtol=1e-6, maxits=60)
            x = K.solve(torch.from_numpy(b[sl].copy()).to(dev)).cpu().numpy()
            r = K.result
            kerr = torch.tensor([rel(x, xo[sl]) if reason_o == 2 else 0.0, float(abs(r["its"] - its_o)), float(r["reason"] != reason_o)],
                                dtype=torch.float64, device=dev)
            dist.all_reduce(kerr, op=dist.ReduceOp.MAX)
            ke = kerr.tolist()
            kgood = ke[0] < 1e-5 and ke[1] <= 1 and ke[2] == 0
            ok = ok and kgood
            if rank == 0:
                print(json.dumps({"check": "slab_ksp", "P": P, "ranks": world, "its": r["its"], "its_oracle": its_o, "x_rel": ke[0], "ok": kgood}), flush=True)
            dist.barrier()
            K.destroy()
        C.destroy()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
