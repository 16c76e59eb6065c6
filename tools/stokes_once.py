"""One StokesFunction + two StokesMatMult at 128^3 (rheology 1, exponent 3, eps 1e-4): the command profiled for the
per-launch time list of the Stokes path (profiles/r01_launches_stokes.csv)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectral_petsc_b200 as sp
P = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
S = sp.Stokes([P] * 3, rheology=1, hardness=1.0, exponent=3.0, regularization=1e-4, gamma0=1.0)
S.set_dirichlet(torch.zeros(S.dv, dtype=torch.float64, device=dev))
S.set_force(torch.zeros(S.g, dtype=torch.float64, device=dev))
xs = torch.from_numpy(0.1 * np.random.default_rng(1).standard_normal(S.g)).to(dev)
x = torch.from_numpy(np.random.default_rng(0).standard_normal(S.g)).to(dev)
y = torch.empty_like(x)
S.function(xs, y)
S.mat_mult(x, y)
torch.cuda.synchronize()
S.mat_mult(x, y)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
