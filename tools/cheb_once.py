"""Three ChebMult applications on a P^3 grid along one axis: the command profiled for the per-P counters (profiles/r02_perP_keys.txt)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectral_petsc_b200 as sp
P, axis = int(sys.argv[1]), int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda:0")
x = torch.from_numpy(np.random.default_rng(0).standard_normal(P ** 3)).to(dev)
y = torch.empty_like(x)
C = sp.Cheb(3, axis, [P] * 3)
for _ in range(3):
    C.mult(x, y)
torch.cuda.synchronize()
print("ok", P, axis, float(y.abs().max()))
