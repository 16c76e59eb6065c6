"""One FGMRES(30) cycle (30 iterations, no PC) on the 128^3 benchmark operator: the command profiled for the per-launch time list
of the Krylov vector work (profiles/r02_launches_ksp.csv)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectral_petsc_b200 as sp
dev = torch.device("cuda:0")
G = sp.Elliptic([128] * 3, gamma=4.0, exponent=2.0)
G.form_function(torch.from_numpy(0.1 * np.random.default_rng(1).standard_normal(G.g)).to(dev))
U = torch.from_numpy(np.random.default_rng(0).standard_normal(G.g)).to(dev)
K = sp.KSP(G.g)
K.set_operators(G)
K.set_tolerances(rtol=1e-30, maxits=30)
K.set_lookahead(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
K.solve(U)
torch.cuda.synchronize()
print("ok", K.result, K.times_ms)
