"""FormFunction + three MatMult_Elliptic at P^3 (default path): the command profiled for the persistent chain kernel's counters."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectral_petsc_b200 as sp
P = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
G = sp.Elliptic([P] * 3, gamma=4.0, exponent=2.0)
G.form_function(torch.from_numpy(0.1 * np.random.default_rng(1).standard_normal(G.g)).to(dev))
U = torch.from_numpy(np.random.default_rng(0).standard_normal(G.g)).to(dev)
V = torch.empty_like(U)
for _ in range(3):
    G.mat_mult(U, V)
torch.cuda.synchronize()
print("ok", P, G.kernel_name(), float(V.abs().max()))
