"""Per-kernel time list of one slab-partitioned StokesMatMult at 128^3 (torch.profiler / CUPTI on every rank, rank 0 prints):
where a step's time goes at N ranks.  torchrun --nproc-per-node N tools/stokes_slab_profile.py"""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectral_petsc_b200 as sp
from spectral_petsc_b200 import dist as spd

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
P = 128
S = sp.Stokes([P] * 3, rheology=1, hardness=1.0, exponent=3.0, regularization=1e-4, gamma0=1.0, rank=rank, nranks=world)
if world > 1:
    spd.attach_peers(S)
S.set_dirichlet(torch.zeros(S.dv, dtype=torch.float64, device=dev))
S.set_force(torch.zeros(S.g, dtype=torch.float64, device=dev))
gen = torch.Generator(device=dev).manual_seed(rank)
xs = 0.1 * torch.randn(S.g, dtype=torch.float64, device=dev, generator=gen)
x = torch.randn(S.g, dtype=torch.float64, device=dev, generator=gen)
y = torch.empty_like(x)
S.function(xs, y)
for _ in range(5):
    S.mat_mult(x, y)
torch.cuda.synchronize(); dist.barrier()
NREP = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(NREP):
        S.mat_mult(x, y)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type.name == "CUDA"]
ev.sort(key=lambda e: e.time_range.start)
if rank == 0:
    t0 = ev[0].time_range.start
    per = len(ev) // NREP
    step = ev[-per:]
    print(json.dumps({"ranks": world, "kernels_per_step": per, "step_span_us": (step[-1].time_range.end - step[0].time_range.start)}))
    for e in step:
        print("%9.1f %8.1f  %s" % (e.time_range.start - step[0].time_range.start, e.time_range.end - e.time_range.start, e.name[:90]))
dist.barrier(); dist.destroy_process_group()
