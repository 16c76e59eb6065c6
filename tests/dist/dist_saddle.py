"""Multi-process slab-partitioned Stokes SOLVE: outer FGMRES(30) + the block-LU saddle-point PC (StokesPCApply0, inner KSPVelocity /
KSPSchur as slab GMRES, cross-rank null-space mean), one rank per GPU, against the same solve on one GPU (rank 0 runs it first).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/dist/dist_saddle.py 32

Prints one JSON line on rank 0; exit code 1 when the iteration counts differ by more than 1 or the solutions by more than 1e-3 (same stopping rule on both: rtol 1e-6 or 150 iterations).
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

KW = dict(vel_max_it=4, schur_max_it=3, svel_preonly=True)


def solve(S, pc, K, rhs, rtol):
    K.set_operators(S, pc=pc)
    K.set_tolerances(rtol=rtol, maxits=150)
    x = K.solve(rhs)
    return x, K.result


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    rheology = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rtol = 1e-6
    dim, d = [P, P, P], 3
    U, U2, dirichlet = sp.stokes_exact_solution(dim, 2)
    dirichlet = dirichlet.reshape(-1)
    mk = lambda r, n: sp.Stokes(dim, rheology=rheology, exponent=2.0 if rheology else 1.0, regularization=0.5 if rheology else 1.0, rank=r, nranks=n)
    gtot = U.size
    ref = torch.zeros(gtot + 2, dtype=torch.float64, device=dev)
    rhs_g = torch.zeros(gtot, dtype=torch.float64, device=dev)
    diag_g = torch.zeros(gtot // (d + 1) * d, dtype=torch.float64, device=dev)  # MatGetDiagonal(MatVVPC): the Jacobi stand-in for PETSc's PC
    if rank == 0:  # the single-GPU solve: the Newton-step system at the exact state (J dx = -F(0-ish)), linearised about U
        S1 = mk(0, 1)
        S1.set_dirichlet(torch.from_numpy(dirichlet.copy()).to(dev))
        S1.set_force(torch.from_numpy(U2).to(dev))
        S1.function(torch.from_numpy(U).to(dev))  # state: eta / deta / strain of the manufactured solution
        rhs1 = torch.from_numpy(np.random.default_rng(0).standard_normal(gtot)).to(dev)
        sp.vec_remove_mean(rhs1, stride=d + 1, offset=d)
        diag_g.copy_(sp.csr_diagonal(*S1.pc_velocity_csr()))
        pc1 = sp.StokesSaddle(S1, 0, velocity_pc=lambda v: v / diag_g, **KW)
        x1, r1 = solve(S1, pc1, sp.KSP(S1.g), rhs1, rtol)
        ref[:gtot] = x1
        ref[gtot] = r1["its"]
        ref[gtot + 1] = r1["reason"]
        rhs_g.copy_(rhs1)
        pc1.destroy()
        S1.destroy()
    dist.broadcast(ref, src=0)
    dist.broadcast(rhs_g, src=0)
    dist.broadcast(diag_g, src=0)
    S = mk(rank, world)
    spd.attach_peers(S)
    S.set_dirichlet(torch.from_numpy(spd.split_dirichlet(dirichlet, dim, world, ncomp=d)[rank].copy()).to(dev))
    S.set_force(torch.from_numpy(spd.split_global(U2, dim, world, ncomp=d + 1)[rank].copy()).to(dev))
    sl = slice((d + 1) * S.goff, (d + 1) * S.goff + S.g)
    S.function(torch.from_numpy(U[sl].copy()).to(dev))
    dloc = diag_g[d * S.goff: d * S.goff + S.gv].clone()
    pc = sp.StokesSaddle(S, 0, velocity_pc=lambda v: v / dloc, **KW)
    spd.attach_peers(pc)
    K = sp.KSP(S.g, rank=rank, nranks=world)
    spd.attach_peers(K)
    torch.cuda.synchronize()
    dist.barrier()
    import time

    t0 = time.perf_counter()
    x, r = solve(S, pc, K, rhs_g[sl].clone(), rtol)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    err = float((x - ref[:gtot][sl]).abs().max())
    e = torch.tensor([err, float(r["its"]), float(S.slab_timeouts()), wall], dtype=torch.float64, device=dev)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    err, its, tmo, wall = e.tolist()
    scale = float(ref[:gtot].abs().max())
    its_ref = int(ref[gtot].item())
    ok = abs(int(its) - its_ref) <= 1 and err <= 1e-3 * scale and tmo == 0 and r["reason"] == int(ref[gtot + 1].item())  # (same stop: converged, or the cap on both)
    if rank == 0:
        print(json.dumps({"check": "slab_saddle_solve", "P": P, "ranks": world, "rheology": rheology, "its": int(its), "its_single_gpu": its_ref, "reason": r["reason"],
                          "inner_its": pc.inner_its, "x_rel": err / scale, "flag_timeouts": tmo, "wall_s": wall, "ok": bool(ok)}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
