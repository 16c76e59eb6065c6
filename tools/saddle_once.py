"""One application of StokesPCApply0 (block LU, -vel_ksp_max_it 4 -schur_ksp_max_it 3 -svel_ksp_type preonly, Jacobi on MatVVPC
standing for the velocity PC so that nothing leaves the device) at P^3: the command profiled for the per-launch time list of the
device-resident saddle-point preconditioner.  Prints the CUDA-event time of one application after a warm-up."""
import os
import sys

import numpy as np
import scipy.sparse as sps
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectral_petsc_b200 as sp  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
S = sp.Stokes([P] * 3, rheology=1, hardness=1.0, exponent=3.0, regularization=1e-4, gamma0=1.0)
S.set_dirichlet(torch.zeros(S.dv, dtype=torch.float64, device=dev))
S.set_force(torch.zeros(S.g, dtype=torch.float64, device=dev))
S.function(torch.from_numpy(0.1 * np.random.default_rng(1).standard_normal(S.g)).to(dev))
rowptr, colidx, vals = [t.cpu().numpy() for t in S.pc_velocity_csr()]
dinv = torch.from_numpy(1.0 / sps.csr_matrix((vals, colidx, rowptr), shape=(S.gv, S.gv)).diagonal()).to(dev)
pc = sp.StokesSaddle(S, 0, velocity_pc=lambda r: dinv * r, vel_max_it=4, schur_max_it=3, svel_preonly=True)
x = torch.from_numpy(np.random.default_rng(0).standard_normal(S.g)).to(dev)
y = torch.empty_like(x)
pc.apply(x, y)
torch.cuda.synchronize()
l0 = sp.launch_count()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
pc.apply(x, y)
b.record()
torch.cuda.synchronize()
print("ok P=%d  StokesPCApply0: %.3f ms, %d launches of this library, inner its %s, max |y| %.3e" % (P, a.elapsed_time(b), sp.launch_count() - l0, pc.inner_its, float(y.abs().max())))
